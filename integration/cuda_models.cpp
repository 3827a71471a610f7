// CUDA-backed subclasses of the reference's models (see cuda_models.h).  Compiled against the reference's
// headers; links libcafe_b200.so through the C ABI in include/cafe_b200.h.
#include "cuda_models.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <numeric>
#include <random>
#include <stdexcept>
#include <unordered_map>

#include "cafe_b200.h"

#include "clade.h"
#include "error_model.h"
#include "gamma.h"
#include "gene_family.h"
#include "io.h"
#include "lambda.h"
#include "matrix_cache.h"
#include "probability.h"
#include "root_distribution.h"
#include "root_equilibrium_distribution.h"
#include "user_data.h"

extern std::mt19937 randomizer_engine;      // the reference's generator (main.cpp)

namespace {

//! What both reference models do before touching the prior (src/base_model.cpp:62-72, src/gamma_core.cpp:182-192).
void initialize_prior(root_equilibrium_distribution* prior, const std::map<int, int>& root_distribution_map, int max_root_family_size)
{
    root_distribution rd;
    if (root_distribution_map.size() > 0)
        rd.vectorize(root_distribution_map);
    else
        rd.vectorize_uniform(max_root_family_size);
    prior->initialize(&rd);
}

bool g_device_branch_probabilities = false;

struct count_row_hash {
    size_t operator()(const std::vector<int>& v) const
    {
        size_t h = 1469598103934665603ull;
        for (int x : v) { h ^= (size_t)x + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); }
        return h;
    }
};

}  // namespace

void cuda_models_use_device_branch_probabilities(bool on) { g_device_branch_probabilities = on; }

// ------------------------------------------------------------------------------------------------------
// cuda_bridge
// ------------------------------------------------------------------------------------------------------

cuda_bridge::cuda_bridge(const clade* p_tree, int max_family_size, int max_root_family_size)
    : _p_tree(p_tree), _mf(max_family_size), _mrf(max_root_family_size)
{
    p_tree->apply_reverse_level_order([this](const clade* c) { _order.push_back(c); });
    std::map<const clade*, int> index;
    for (size_t i = 0; i < _order.size(); ++i) index[_order[i]] = (int)i;
    const int nn = (int)_order.size();
    _parent.assign(nn, -1);
    _leaf_col.assign(nn, -1);
    _lambda_index.resize(nn);
    _branch.resize(nn);
    _child_offset.assign(nn + 1, 0);
    for (int v = 0; v < nn; ++v) {
        const clade* c = _order[v];
        _parent[v] = c->is_root() ? -1 : index.at(c->get_parent());
        _branch[v] = c->get_branch_length();
        _lambda_index[v] = v;                       // one lambda slot per node: the value comes from lambda::get_value_for_clade
        if (c->is_leaf()) {
            _leaf_col[v] = (int)_leaves.size();
            _leaves.push_back(c);
        }
        else {
            _internal.push_back(c);
        }
        c->apply_to_descendants([&](const clade* d) { _child_list.push_back(index.at(d)); });     // Newick order
        _child_offset[v + 1] = (int)_child_list.size();
    }
}

cuda_bridge::~cuda_bridge()
{
    release();
}

std::vector<double> cuda_bridge::host_seconds() const
{
    double lib[3] = {0.0, 0.0, 0.0};
    if (_ctx) cafe_b200_host_seconds(_ctx, lib);
    return {_bind_seconds, _eval_call_seconds, _lib_seconds_before[0] + lib[0], _lib_seconds_before[1] + lib[1], _lib_seconds_before[2] + lib[2]};
}

void cuda_bridge::release()
{
    if (_ctx) {
        double lib[3] = {0.0, 0.0, 0.0};
        cafe_b200_host_seconds(_ctx, lib);
        for (int i = 0; i < 3; ++i) _lib_seconds_before[i] += lib[i];
    }
    if (_ctx) cafe_b200_destroy(_ctx);
    _ctx = nullptr;
    cafe_b200_free_pinned(_h_family_lnl);
    cafe_b200_free_pinned(_h_cat_lk);
    _h_family_lnl = _h_cat_lk = nullptr;
    _cat_lk_capacity = 0;
    _bound = nullptr;
    _bound_size = 0;
}

void cuda_bridge::check(int rc, const char* what) const
{
    if (rc != CAFE_B200_OK)
        throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + cafe_b200_last_error(_ctx));
}

std::vector<int> cuda_bridge::devices_from_environment()
{
    std::vector<int> devices;
    if (const char* e = getenv("CAFE_B200_DEVICES")) {
        std::string list = e;
        if (list == "all") {
            for (int d = 0; d < cafe_b200_device_count(); ++d) devices.push_back(d);
        }
        else {
            size_t pos = 0;
            while (pos < list.size()) {
                size_t comma = list.find(',', pos);
                if (comma == std::string::npos) comma = list.size();
                if (comma > pos) devices.push_back(atoi(list.substr(pos, comma - pos).c_str()));
                pos = comma + 1;
            }
        }
    }
    if (devices.empty()) {
        int device = 0;
        if (const char* e = getenv("CAFE_B200_DEVICE")) device = atoi(e);
        devices.push_back(device);
    }
    return devices;
}

int cuda_bridge::device_count() const { return _ctx ? cafe_b200_n_devices(_ctx) : 0; }

// Counts of a spread of rows, about 256 look-ups in all (a few microseconds; a full pass costs F x leaves string-keyed
// map look-ups, milliseconds to seconds): enough to notice a bound vector that was edited in place.
size_t cuda_bridge::fingerprint(const std::vector<gene_family>& families) const
{
    size_t h = 1469598103934665603ull;
    const size_t n = families.size(), rows = std::max<size_t>(2, 256 / std::max<size_t>(1, _leaves.size()));
    const size_t step = std::max<size_t>(1, n / rows);
    for (size_t i = 0; i < n; i += step)
        for (const clade* leaf : _leaves) h = (h ^ (size_t)families[i].get_species_size(leaf->get_taxon_name())) * 1099511628211ull;
    return h;
}

void cuda_bridge::bind(const std::vector<gene_family>& families)
{
    if (_ctx && _bound == &families && _bound_size == families.size() && _bound_fingerprint == fingerprint(families)) return;
    const auto t0 = std::chrono::steady_clock::now();
    const size_t nl = _leaves.size();
    std::vector<int> rows(families.size() * nl);
    std::vector<std::string> names(nl);
    for (size_t l = 0; l < nl; ++l) names[l] = _leaves[l]->get_taxon_name();
    // F x leaves look-ups in each family's string-keyed map: spread over the host threads the reference's own loops use
    const long n_fam = (long)families.size();
#pragma omp parallel for schedule(static) if (n_fam * (long)nl > 100000)
    for (long i = 0; i < n_fam; ++i)
        for (size_t l = 0; l < nl; ++l) rows[(size_t)i * nl + l] = families[i].get_species_size(names[l]);
    bind_rows(rows, families.size());
    _bound = &families;
    _bound_size = families.size();
    _bound_fingerprint = fingerprint(families);
    _bind_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

void cuda_bridge::bind_rows(const std::vector<int>& rows, size_t n_rows)
{
    release();
    ++_generation;

    // Identical count rows are evaluated once.  The reference does this for the base model only
    // (build_reference_list, src/base_model.cpp:27-51, O(F^2)); identical inputs give identical outputs for
    // the gamma model as well, so de-duplicating there changes no number.
    const size_t nl = _leaves.size();
    std::unordered_map<std::vector<int>, size_t, count_row_hash> seen;
    std::vector<int32_t> counts;
    _unique_of.resize(n_rows);
    _max_count = 0;
    int min_count = 0;
    std::vector<int> row(nl);
    for (size_t i = 0; i < n_rows; ++i) {
        row.assign(rows.begin() + i * nl, rows.begin() + (i + 1) * nl);
        auto it = seen.find(row);
        if (it == seen.end()) {
            it = seen.emplace(row, seen.size()).first;
            counts.insert(counts.end(), row.begin(), row.end());
            for (int x : row) { _max_count = std::max(_max_count, x); min_count = std::min(min_count, x); }
        }
        _unique_of[i] = it->second;
    }
    _n_unique = seen.size();

    cafe_b200_tree t;
    t.n_nodes = (int)_order.size();
    t.parent = _parent.data();
    t.child_offset = _child_offset.data();
    t.child_list = _child_list.data();
    t.leaf_col = _leaf_col.data();
    t.branch = _branch.data();
    t.lambda_index = _lambda_index.data();
    const std::vector<int> devices = devices_from_environment();
    // counts travel as one byte each when they fit (the device stores them that way for max_family_size <= 255)
    std::vector<uint8_t> narrow;
    const void* data = counts.data();
    int width = 4;
    if (min_count >= 0 && _max_count <= 255) {
        narrow.assign(counts.begin(), counts.end());
        data = narrow.data();
        width = 1;
    }
    int rc = cafe_b200_create_multi(&_ctx, &t, data, width, (int64_t)_n_unique, (int)nl, _mf, _mrf, devices.data(), (int)devices.size());
    if (rc != CAFE_B200_OK) {
        _ctx = nullptr;
        throw std::runtime_error(std::string("cafe_b200_create failed (") + std::to_string(rc) + "): " + cafe_b200_last_error(nullptr));
    }
    _h_family_lnl = static_cast<double*>(cafe_b200_alloc_pinned(std::max<size_t>(1, _n_unique) * sizeof(double)));
    if (!_h_family_lnl) throw std::runtime_error("cafe_b200_alloc_pinned failed");
    _failed_idx.assign(std::max<size_t>(1, _n_unique), 0);
}

void cuda_bridge::set_error_model(const error_model* p_error_model)
{
    if (!p_error_model) {
        check(cafe_b200_set_error_model(_ctx, nullptr, 0, 0), "cafe_b200_set_error_model");
        return;
    }
    // rows for every observed count that occurs: error_model::get_probs(count)  (src/probability.cpp:182)
    const int nd = (int)p_error_model->n_deviations();
    const int rows = _max_count + 1;
    std::vector<double> table((size_t)rows * nd);
    for (int s = 0; s < rows; ++s) {
        std::vector<double> probs = p_error_model->get_probs(s);
        for (int d = 0; d < nd; ++d) table[(size_t)s * nd + d] = probs[d];
    }
    check(cafe_b200_set_error_model(_ctx, table.data(), rows, nd), "cafe_b200_set_error_model");
}

std::vector<double> cuda_bridge::lambda_table(const lambda* p_lambda, const std::vector<double>& multipliers) const
{
    const size_t nn = _order.size();
    std::vector<double> out(multipliers.size() * nn);
    for (size_t k = 0; k < multipliers.size(); ++k) {
        std::unique_ptr<lambda> ml(p_lambda->multiply(multipliers[k]));            // src/core.cpp:135
        for (size_t v = 0; v < nn; ++v)
            out[k * nn + v] = _order[v]->is_root() ? ml->get_value_for_clade(_order[0]) : ml->get_value_for_clade(_order[v]);
    }
    return out;
}

std::vector<double> cuda_bridge::prior_table(const root_equilibrium_distribution* prior, int n)
{
    std::vector<double> out(n);
    for (int j = 0; j < n; ++j) out[j] = (double)prior->compute(j);
    return out;
}

long cuda_bridge::evaluate(const std::vector<double>& lambdas, const std::vector<double>& cat_probs, const std::vector<double>& prior, int mode,
                           std::vector<char>& failed)
{
    const int k = (int)cat_probs.size();
    double neg_lnl = 0.0;
    int64_t n_failed = 0;
    static_assert(sizeof(long long) == sizeof(int64_t), "64-bit indices");
    // only the per-family lnL comes back with every evaluation (the caller sums it in family order); the category
    // likelihoods stay on the device until somebody asks (category_likelihoods())
    const auto t0 = std::chrono::steady_clock::now();
    check(cafe_b200_eval(_ctx, lambdas.data(), (int)_order.size(), cat_probs.data(), k, prior.data(), mode, &neg_lnl, _h_family_lnl, nullptr,
                         &n_failed, reinterpret_cast<int64_t*>(_failed_idx.data()), (int64_t)_failed_idx.size()),
          "cafe_b200_eval");
    _eval_call_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    _last_k = mode == CAFE_B200_GAMMA_LINSUM ? k : 0;
    _cat_lk_fetched = false;
    ++_evaluations;
    double ms[4] = {0, 0, 0, 0};
    if (cafe_b200_last_timings(_ctx, ms) == CAFE_B200_OK) _device_seconds += (ms[0] + ms[1] + ms[2]) * 1e-3;
    failed.clear();
    if (n_failed > 0) {
        failed.assign(_n_unique, 0);
        for (int64_t i = 0; i < n_failed; ++i) failed[_failed_idx[i]] = 1;
    }
    return (long)n_failed;
}

const double* cuda_bridge::category_likelihoods()
{
    if (_last_k <= 0) throw std::runtime_error("category_likelihoods: the last evaluation was not a gamma evaluation");
    if (!_cat_lk_fetched) {
        const size_t need = std::max<size_t>(1, _n_unique * _last_k);
        if (need > _cat_lk_capacity) {
            cafe_b200_free_pinned(_h_cat_lk);
            _h_cat_lk = static_cast<double*>(cafe_b200_alloc_pinned(need * sizeof(double)));
            if (!_h_cat_lk) throw std::runtime_error("cafe_b200_alloc_pinned failed");
            _cat_lk_capacity = need;
        }
        check(cafe_b200_fetch_category_likelihoods(_ctx, _last_k, _h_cat_lk), "cafe_b200_fetch_category_likelihoods");
        _cat_lk_fetched = true;
    }
    return _h_cat_lk;
}

std::vector<double> cuda_bridge::root_max(const std::vector<double>& lambdas)
{
    std::vector<double> out(_n_unique, 0.0);
    check(cafe_b200_root_max(_ctx, lambdas.data(), (int)_order.size(), out.data()), "cafe_b200_root_max");
    return out;
}

std::vector<double> cuda_bridge::branch_probabilities(const std::vector<double>& lambdas, const std::vector<int>& node_sizes,
                                                      const std::vector<unsigned char>& selected)
{
    // the context must hold exactly one row per family here (bind_rows de-duplicates count rows, but two families with
    // the same counts still share their reconstruction, so the result is expanded by the caller through unique_of)
    std::vector<double> out(_n_unique * _order.size(), -1.0);
    static_assert(sizeof(int) == sizeof(int32_t), "int is 32 bits");
    check(cafe_b200_branch_probabilities(_ctx, lambdas.data(), (int)_order.size(), reinterpret_cast<const int32_t*>(node_sizes.data()),
                                         selected.empty() ? nullptr : selected.data(), out.data()),
          "cafe_b200_branch_probabilities");
    return out;
}

void cuda_bridge::reconstruct(const std::vector<double>& lambdas, int n_categories, const std::vector<double>& prior_by_size, std::vector<int>& states)
{
    states.assign(_n_unique * n_categories * _internal.size(), 0);
    static_assert(sizeof(int) == sizeof(int32_t), "int is 32 bits");
    check(cafe_b200_reconstruct(_ctx, lambdas.data(), (int)_order.size(), n_categories, prior_by_size.data(), reinterpret_cast<int32_t*>(states.data())),
          "cafe_b200_reconstruct");
}

// ------------------------------------------------------------------------------------------------------
// cuda_base_model
// ------------------------------------------------------------------------------------------------------

cuda_base_model::cuda_base_model(lambda* p_lambda, const clade* p_tree, const std::vector<gene_family>* p_gene_families, int max_family_size,
                                 int max_root_family_size, error_model* p_error_model)
    : base_model(p_lambda, p_tree, p_gene_families, max_family_size, max_root_family_size, p_error_model),
      _bridge(p_tree, max_family_size, max_root_family_size)
{
}

double cuda_base_model::infer_family_likelihoods(root_equilibrium_distribution* prior, const std::map<int, int>& root_distribution_map, const lambda* p_lambda)
{
    _monitor.Event_InferenceAttempt_Started();
    if (!_p_lambda->is_valid()) {                                                   // src/base_model.cpp:56-60
        _monitor.Event_InferenceAttempt_InvalidValues();
        return -log(0);
    }
    initialize_prior(prior, root_distribution_map, _max_root_family_size);

    _bridge.bind(*_p_gene_families);
    _bridge.set_error_model(_p_error_model);                                       // the epsilon optimiser edits it in place between calls
    std::vector<char> failed;
    _bridge.evaluate(_bridge.lambda_table(_p_lambda, {1.0}), {1.0}, cuda_bridge::prior_table(prior, _max_root_family_size),
                     CAFE_B200_BASE_LOGMAX, failed);

    // summed on the host in family order, as the reference does (src/base_model.cpp:107)
    const size_t F = _p_gene_families->size();
    const double* family_lnl = _bridge.family_lnl();
    double sum = 0.0;
    for (size_t i = 0; i < F; ++i) sum += family_lnl[_bridge.unique_of(i)];
    const double final_likelihood = -sum;
    _results_stale = true;
    _generation_of_results = _bridge.generation();
    _monitor.Event_InferenceAttempt_Complete(final_likelihood);
    return final_likelihood;
}

void cuda_base_model::materialize_results()
{
    if (!_results_stale) return;
    if (_generation_of_results != _bridge.generation())
        throw std::runtime_error("cuda_base_model: the device context was rebuilt after the last evaluation; evaluate again before reading results");
    const size_t F = _p_gene_families->size();
    const double* family_lnl = _bridge.family_lnl();
    results.resize(F);
    for (size_t i = 0; i < F; ++i)
        results[i] = family_info_stash(_p_gene_families->at(i).id(), 0.0, 0.0, 0.0, family_lnl[_bridge.unique_of(i)], false);      // src/base_model.cpp:105
    _results_stale = false;
}

void cuda_base_model::write_family_likelihoods(std::ostream& ost)
{
    materialize_results();
    base_model::write_family_likelihoods(ost);
}

reconstruction* cuda_base_model::reconstruct_ancestral_states(const std::vector<gene_family>& families, matrix_cache* p_calc, root_equilibrium_distribution* p_prior)
{
    _monitor.Event_Reconstruction_Started("Base");
    auto result = new base_model_reconstruction();
    // callers read transition matrices from *p_calc afterwards (compute_viterbi_sum, src/execute.cpp:158-170) unless
    // they take the branch probabilities from the device
    if (!g_device_branch_probabilities) p_calc->precalculate_matrices(get_lambda_values(_p_lambda), _p_tree->get_branch_lengths());

    materialize_results();                                                         // before the bridge is re-bound to `families`
    _bridge.bind(families);
    const int lim = std::min(_max_family_size, _max_root_family_size) + 1;         // src/gene_family_reconstructor.cpp:41-56
    std::vector<int> states;
    _bridge.reconstruct(_bridge.lambda_table(_p_lambda, {1.0}), 1, cuda_bridge::prior_table(p_prior, lim), states);
    const auto& internal = _bridge.internal_nodes();
    for (size_t i = 0; i < families.size(); ++i) {
        clademap<int>& r = result->_reconstructions[families[i].id()];
        const int* row = &states[_bridge.unique_of(i) * internal.size()];
        for (size_t n = 0; n < internal.size(); ++n) r[internal[n]] = row[n];
    }
    _monitor.Event_Reconstruction_Complete();
    return result;
}

// ------------------------------------------------------------------------------------------------------
// cuda_gamma_model
// ------------------------------------------------------------------------------------------------------

cuda_gamma_model::cuda_gamma_model(lambda* p_lambda, clade* p_tree, std::vector<gene_family>* p_gene_families, int max_family_size,
                                   int max_root_family_size, int n_gamma_cats, double fixed_alpha, error_model* p_error_model)
    : gamma_model(p_lambda, p_tree, p_gene_families, max_family_size, max_root_family_size, n_gamma_cats, fixed_alpha, p_error_model),
      _bridge(p_tree, max_family_size, max_root_family_size), _explicit_categories(false)
{
}

cuda_gamma_model::cuda_gamma_model(lambda* p_lambda, clade* p_tree, std::vector<gene_family>* p_gene_families, int max_family_size,
                                   int max_root_family_size, std::vector<double> gamma_categories, std::vector<double> multipliers,
                                   error_model* p_error_model)
    : gamma_model(p_lambda, p_tree, p_gene_families, max_family_size, max_root_family_size, gamma_categories, multipliers, p_error_model),
      _bridge(p_tree, max_family_size, max_root_family_size), _explicit_categories(true), _explicit_cat_probs(gamma_categories)
{
}

std::vector<double> cuda_gamma_model::cat_probs() const
{
    if (_explicit_categories) return _explicit_cat_probs;
    // gamma_model keeps _gamma_cat_probs private; they are a pure function of (k, alpha): src/gamma_core.cpp:58-64
    const size_t k = get_gamma_cat_probs_count();
    std::vector<double> freq(k), rate(k);
    if (k > 1) get_gamma(freq, rate, get_alpha());
    return freq;
}

double cuda_gamma_model::infer_family_likelihoods(root_equilibrium_distribution* prior, const std::map<int, int>& root_distribution_map, const lambda* p_lambda)
{
    _monitor.Event_InferenceAttempt_Started();
    results.clear();
    _results_stale = _cat_lk_stale = false;
    if (!can_infer()) {                                                             // src/gamma_core.cpp:175-179
        _monitor.Event_InferenceAttempt_InvalidValues();
        return -log(0);
    }
    initialize_prior(prior, root_distribution_map, _max_root_family_size);

    _last_multipliers = get_lambda_multipliers();
    _last_probs = cat_probs();
    _bridge.bind(*_p_gene_families);
    _bridge.set_error_model(_p_error_model);
    std::vector<char> failed;
    const long n_failed = _bridge.evaluate(_bridge.lambda_table(p_lambda, _last_multipliers), _last_probs,
                                           cuda_bridge::prior_table(prior, _max_root_family_size), CAFE_B200_GAMMA_LINSUM, failed);

    const size_t F = _p_gene_families->size();
    _cat_lk.clear();
    _last_failed = n_failed > 0;
    if (n_failed > 0) {                                                             // src/gamma_core.cpp:227-236
        for (size_t i = 0; i < F; ++i)
            if (failed[_bridge.unique_of(i)]) _monitor.Event_InferenceAttempt_Saturation(_p_gene_families->at(i).id());
        return -log(0);
    }
    // lnL_i = log(sum_k cat_lk[i][k]) comes from the device per unique family (src/gamma_core.cpp:207,218); the sum
    // over families is formed on the host in family order, as the reference forms it (src/gamma_core.cpp:244)
    const double* family_lnl = _bridge.family_lnl();
    double sum = 0.0;
    for (size_t i = 0; i < F; ++i) sum += family_lnl[_bridge.unique_of(i)];
    const double final_likelihood = -sum;
    _results_stale = _cat_lk_stale = true;
    _generation_of_results = _bridge.generation();
    _monitor.Event_InferenceAttempt_Complete(final_likelihood);
    return final_likelihood;
}

const std::vector<std::vector<double>>& cuda_gamma_model::category_likelihoods()
{
    if (_cat_lk_stale) {
        if (_generation_of_results != _bridge.generation())
            throw std::runtime_error("cuda_gamma_model: the device context was rebuilt after the last evaluation; evaluate again before reading results");
        const size_t F = _p_gene_families->size(), k = _last_probs.size();
        const double* cat_lk = _bridge.category_likelihoods();
        _cat_lk.assign(F, std::vector<double>());
        for (size_t i = 0; i < F; ++i) {
            const double* cl = cat_lk + _bridge.unique_of(i) * k;
            _cat_lk[i].assign(cl, cl + k);
        }
        _cat_lk_stale = false;
    }
    return _cat_lk;
}

void cuda_gamma_model::materialize_results()
{
    if (!_results_stale) return;
    const std::vector<std::vector<double>>& all = category_likelihoods();
    const size_t F = _p_gene_families->size(), k = _last_probs.size();
    results.clear();
    results.reserve(F * k);
    for (size_t i = 0; i < F; ++i) {
        const std::vector<double>& cl = all[i];
        const double family_likelihood = std::accumulate(cl.begin(), cl.end(), 0.0);     // src/gamma_core.cpp:207
        // posterior: the reference multiplies by the category probability a second time (src/gamma_core.cpp:97-109)
        double denominator = 0.0;
        for (size_t c = 0; c < k; ++c) denominator += cl[c] * _last_probs[c];
        for (size_t c = 0; c < k; ++c) {
            const double posterior = cl[c] * _last_probs[c] / denominator;
            results.push_back(family_info_stash(_p_gene_families->at(i).id(), _last_multipliers[c], cl[c], family_likelihood, posterior, posterior > 0.95));
        }
    }
    _results_stale = false;
}

void cuda_gamma_model::write_family_likelihoods(std::ostream& ost)
{
    materialize_results();
    gamma_model::write_family_likelihoods(ost);
}

reconstruction* cuda_gamma_model::reconstruct_ancestral_states(const std::vector<gene_family>& families, matrix_cache* p_calc, root_equilibrium_distribution* p_prior)
{
    _monitor.Event_Reconstruction_Started("Gamma");
    const std::vector<double> multipliers = get_lambda_multipliers();
    const std::vector<double> probs = cat_probs();
    const size_t k = multipliers.size();

    if (!g_device_branch_probabilities) {
        std::vector<double> all;                                                    // src/gamma_core.cpp:305-315
        for (double multiplier : multipliers)
            for (double lam : get_lambda_values(_p_lambda)) all.push_back(lam * multiplier);
        p_calc->precalculate_matrices(all, _p_tree->get_branch_lengths());
    }

    // the last evaluation's outputs leave the device before the bridge is re-bound to `families`
    if (!_last_failed && (_results_stale || _cat_lk_stale)) { category_likelihoods(); materialize_results(); }
    _bridge.bind(families);
    const int lim = std::min(_max_family_size, _max_root_family_size) + 1;
    std::vector<int> states;
    _bridge.reconstruct(_bridge.lambda_table(_p_lambda, multipliers), (int)k, cuda_bridge::prior_table(p_prior, lim), states);

    gamma_model_reconstruction* result = new gamma_model_reconstruction(multipliers);
    const auto& internal = _bridge.internal_nodes();
    for (size_t i = 0; i < families.size(); ++i) {
        auto& rec = result->_reconstructions[families[i].id()];
        if (i < _cat_lk.size()) rec._category_likelihoods = _cat_lk[i];
        rec.category_reconstruction.resize(k);
        const int* row = &states[_bridge.unique_of(i) * k * internal.size()];
        for (size_t c = 0; c < k; ++c)
            for (size_t n = 0; n < internal.size(); ++n) rec.category_reconstruction[c][internal[n]] = row[c * internal.size() + n];
        rec.reconstruction = get_weighted_averages(rec.category_reconstruction, probs);   // src/gamma_core.cpp:338-342
    }
    _monitor.Event_Reconstruction_Complete();
    return result;
}

// ------------------------------------------------------------------------------------------------------

// ------------------------------------------------------------------------------------------------------
// p-values
// ------------------------------------------------------------------------------------------------------

std::vector<double> compute_pvalues_cuda(const clade* p_tree, const std::vector<gene_family>& families, const lambda* p_lambda,
                                         const matrix_cache& cache, int number_of_simulations, int max_family_size, int max_root_family_size)
{
    const int mx = max_family_size, mxr = max_root_family_size, nsim = number_of_simulations;

    // (1) simulate as get_random_probabilities does, root size by root size (src/probability.cpp:279-298), drawing from
    //     randomizer_engine in the reference's order with the reference's distributions, so every simulated family is
    //     the reference's.  What set_weighted_random_family_size (src/probability.cpp:320-351) rebuilds per call — the
    //     weight vector of row `parent size` and its std::discrete_distribution — depends only on (edge, parent size)
    //     and is built once here; a distribution object draws the same values however often it is reused.
    std::vector<const clade*> prefix;
    p_tree->apply_prefix_order([&](const clade* c) { prefix.push_back(c); });
    std::map<const clade*, int> pos;
    for (size_t i = 0; i < prefix.size(); ++i) pos[prefix[i]] = (int)i;
    struct edge_sampler {
        int parent = -1;
        int leaf_col = -1;
        bool saturated = false;
        const matrix* probabilities = nullptr;
        std::vector<std::unique_ptr<std::discrete_distribution<int>>> by_parent_size;
    };
    cuda_bridge sim_bridge(p_tree, mx, mxr);
    const std::vector<const clade*>& leaves = sim_bridge.leaf_nodes();
    std::vector<edge_sampler> edge(prefix.size());
    for (size_t i = 0; i < prefix.size(); ++i) {
        const clade* c = prefix[i];
        if (c->is_root()) continue;
        edge[i].parent = pos.at(c->get_parent());
        const double lam = p_lambda->get_value_for_clade(c), t = c->get_branch_length();
        edge[i].probabilities = cache.get_matrix(t, lam);
        edge[i].saturated = cache.is_saturated(t, lam);
        edge[i].by_parent_size.resize(edge[i].probabilities->size());
        if (c->is_leaf()) edge[i].leaf_col = (int)(std::find(leaves.begin(), leaves.end(), c) - leaves.begin());
    }
    const size_t nl = leaves.size();
    std::vector<int> sim_rows((size_t)mxr * nsim * nl, 0);
    std::vector<int> sizes(prefix.size());
    std::vector<double> v(mx);
    for (int root_size = 0; root_size < mxr; ++root_size) {
        for (int i = 0; i < nsim; ++i) {
            int* row = &sim_rows[((size_t)root_size * nsim + i) * nl];
            for (size_t n = 0; n < prefix.size(); ++n) {
                edge_sampler& e = edge[n];
                if (e.parent < 0) { sizes[n] = root_size; continue; }
                const int parent_size = sizes[e.parent];
                int c = 0;
                if (parent_size > 0) {
                    if (e.saturated) {
                        std::uniform_int_distribution<int> distribution(0, mx - 1);      // drawn and overwritten, as in the reference
                        c = distribution(randomizer_engine);
                    }
                    auto& dist = e.by_parent_size[parent_size];
                    if (!dist) {
                        for (int k = 0; k < mx; ++k) v[k] = e.probabilities->get(parent_size, k);
                        dist.reset(new std::discrete_distribution<int>(v.begin(), v.end()));
                    }
                    c = (*dist)(randomizer_engine);
                }
                sizes[n] = c;
                if (e.leaf_col >= 0) row[e.leaf_col] = c;
            }
        }
    }

    // (2) likelihood of a family = max of its root vector (src/probability.cpp:308, 399): pruning kernel
    const std::vector<double> lambdas = sim_bridge.lambda_table(p_lambda, std::vector<double>{1.0});
    sim_bridge.bind_rows(sim_rows, (size_t)mxr * nsim);
    std::vector<double> cond((size_t)mxr * nsim);
    {
        const std::vector<double> unique = sim_bridge.root_max(lambdas);
        for (size_t i = 0; i < cond.size(); ++i) cond[i] = unique[sim_bridge.unique_of(i)];
    }
    std::vector<double> observed(families.size());
    {
        cuda_bridge bridge(p_tree, mx, mxr);
        bridge.bind(families);
        const std::vector<double> unique = bridge.root_max(lambdas);
        for (size_t i = 0; i < observed.size(); ++i) observed[i] = unique[bridge.unique_of(i)];
    }

    // (3) sort each conditional distribution, upper_bound, max over root sizes (src/probability.cpp:310, 379-409)
    std::vector<double> result(families.size());
    const int device = cuda_bridge::devices_from_environment()[0];
    int rc = cafe_b200_pvalues(device, cond.data(), mxr, nsim, observed.data(), (int64_t)observed.size(), result.data());
    if (rc != CAFE_B200_OK) throw std::runtime_error(std::string("cafe_b200_pvalues failed (") + std::to_string(rc) + "): " + cafe_b200_last_error(nullptr));
    return result;
}

branch_probabilities compute_branch_probabilities_cuda(const clade* p_tree, const std::vector<gene_family>& families, const reconstruction* rec,
                                                       const std::vector<double>& pvalues, double test_pvalue, const lambda* p_lambda,
                                                       int max_family_size, int max_root_family_size)
{
    cuda_bridge bridge(p_tree, max_family_size, max_root_family_size);
    bridge.bind(families);
    const std::vector<const clade*>& order = bridge.nodes();
    const size_t nn = order.size(), nu = bridge.unique_count();
    // one row of reconstructed sizes per UNIQUE count row (identical families have identical reconstructions)
    std::vector<int> sizes(nu * nn, 0);
    std::vector<unsigned char> selected(nu, 0);
    for (size_t i = 0; i < families.size(); ++i) {
        const size_t u = bridge.unique_of(i);
        if (!(pvalues[i] < test_pvalue)) continue;                   // src/execute.cpp:169
        if (selected[u]) continue;
        selected[u] = 1;
        for (size_t v = 0; v < nn; ++v) sizes[u * nn + v] = rec->reconstructed_size(families[i], order[v]);
    }
    const std::vector<double> probs = bridge.branch_probabilities(bridge.lambda_table(p_lambda, std::vector<double>{1.0}), sizes, selected);
    branch_probabilities result;
    for (size_t i = 0; i < families.size(); ++i) {
        if (!(pvalues[i] < test_pvalue)) continue;
        const size_t u = bridge.unique_of(i);
        for (size_t v = 0; v < nn; ++v) {
            const double p = probs[u * nn + v];
            result.set(families[i], order[v], p < 0 ? branch_probabilities::invalid() : branch_probabilities::branch_probability(p));
        }
    }
    return result;
}

std::vector<model*> build_cuda_models(const input_parameters& user_input, user_data& user_data)
{
    model* p_model = NULL;
    std::vector<gene_family>* p_gene_families = &user_data.gene_families;
    if (user_input.is_simulating) p_gene_families = NULL;

    if (user_input.fixed_alpha > 0 || user_input.n_gamma_cats > 1) {
        p_model = new cuda_gamma_model(user_data.p_lambda, user_data.p_tree, &user_data.gene_families, user_data.max_family_size,
                                       user_data.max_root_family_size, user_input.n_gamma_cats, user_input.fixed_alpha, user_data.p_error_model);
    }
    else {
        error_model* p_error_model = user_data.p_error_model;
        if (user_input.use_error_model && !p_error_model) {                         // src/core.cpp:37-42
            p_error_model = new error_model();
            p_error_model->set_probabilities(0, {0, .95, 0.05});
            p_error_model->set_probabilities(user_data.max_family_size, {0.05, .9, 0.05});
        }
        p_model = new cuda_base_model(user_data.p_lambda, user_data.p_tree, p_gene_families, user_data.max_family_size,
                                      user_data.max_root_family_size, p_error_model);
    }
    return std::vector<model*>{p_model};
}
