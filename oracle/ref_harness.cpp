// TEST INFRASTRUCTURE — not part of the shipped product.
//
// Driver around the UNMODIFIED reference (Han9527/CAFExp) objects, which oracle/Makefile compiles
// from the sources where they lie under /root/reference into oracle/_ref/.  It is used to
//   (1) pin the C restatement in oracle/cafe_oracle.c (golden vectors under tests/golden/ are
//       produced by scripts/make_golden.py running this binary), and
//   (2) time the reference's own CPU path on the GPU box's host cores (bench.py --impl reference,
//       cpu_baseline.kind == "reference").
// Nothing in cafexp_b200/ links or executes it.
//
// Every sub-command prints one JSON object on stdout (doubles with 17 significant digits) and
// may write raw little-endian arrays to the file given with --dump.
//
// Reference entry points exercised (file:line in /root/reference):
//   the_probability_of_going_from_parent_fam_size_to_c      src/probability.cpp:147
//   matrix_cache::precalculate_matrices / get_matrix        src/matrix_cache.cpp:121 / :80
//   inference_prune                                         src/core.cpp:133
//   base_model::infer_family_likelihoods                    src/base_model.cpp:53
//   gamma_model::infer_family_likelihoods                   src/gamma_core.cpp:169
//   model::reconstruct_ancestral_states                     src/base_model.cpp:145, src/gamma_core.cpp:301
//   optimizer::optimize                                     src/optimizer.cpp:539
//   get_gamma                                               src/gamma.cpp:225
//   compute_pvalues / set_weighted_random_family_size       src/probability.cpp:411 / :320
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <numeric>
#include <random>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include <deque>
#include <queue>
#include <stack>
#include <omp.h>

// The reference keeps per-family results in protected/private members; the harness reads them
// directly instead of re-parsing the 6-digit text reports.
#define private public
#define protected public
#include "io.h"
#include "clade.h"
#include "core.h"
#include "base_model.h"
#include "gamma_core.h"
#define CAFEXP_GAMMA_CORE_H_INCLUDED
#include "gamma.h"
#include "lambda.h"
#include "matrix_cache.h"
#include "probability.h"
#include "error_model.h"
#include "gene_family.h"
#include "user_data.h"
#include "root_equilibrium_distribution.h"
#include "root_distribution.h"
#include "optimizer.h"
#include "optimizer_scorer.h"
#include "gene_family_reconstructor.h"
#undef private
#undef protected
#ifdef WITH_CUDA_MODELS
// Second build of this driver (oracle/_ref/ref_harness_cuda): the same unmodified reference host code —
// optimizer, scorers, readers — with the model objects swapped for the CUDA-backed subclasses of
// integration/cuda_models.cpp when --cuda 1 is given.  Used by the GPU tests for fit parity.
#include "cuda_models.h"
#endif

std::mt19937 randomizer_engine(10);   // the reference's main.cpp defines this global; we replace main.cpp

void init_lgamma_cache();

namespace {

struct args_t {
    std::map<std::string, std::string> kv;
    bool has(const std::string& k) const { return kv.count(k) != 0; }
    std::string str(const std::string& k, const std::string& d = "") const { auto it = kv.find(k); return it == kv.end() ? d : it->second; }
    double num(const std::string& k, double d) const { return has(k) ? atof(kv.at(k).c_str()) : d; }
    long integer(const std::string& k, long d) const { return has(k) ? atol(kv.at(k).c_str()) : d; }
};

args_t parse(int argc, char** argv, int first)
{
    args_t a;
    for (int i = first; i < argc; ++i) {
        std::string k = argv[i];
        if (k.rfind("--", 0) != 0) throw std::runtime_error("bad argument " + k);
        k = k.substr(2);
        std::string v = "1";
        if (i + 1 < argc && std::string(argv[i + 1]).rfind("--", 0) != 0) v = argv[++i];
        a.kv[k] = v;
    }
    return a;
}

void jnum(double v)
{
    if (std::isinf(v)) printf(v > 0 ? "\"inf\"" : "\"-inf\"");
    else if (std::isnan(v)) printf("\"nan\"");
    else printf("%.17g", v);
}

void jarr(const char* name, const std::vector<double>& v)
{
    printf("\"%s\": [", name);
    for (size_t i = 0; i < v.size(); ++i) { if (i) printf(", "); jnum(v[i]); }
    printf("]");
}

struct dumper {
    FILE* f = nullptr;
    explicit dumper(const std::string& path) { if (!path.empty()) { f = fopen(path.c_str(), "wb"); if (!f) throw std::runtime_error("cannot open " + path); } }
    ~dumper() { if (f) fclose(f); }
    void doubles(const double* p, size_t n) { if (f) fwrite(p, sizeof(double), n, f); }
    void ints(const int* p, size_t n) { if (f) fwrite(p, sizeof(int), n, f); }
};

// Everything cafexp() builds before act->execute(), src/cafexp.cpp:175-204, driven from --options.
struct setup {
    input_parameters in;
    user_data data;
    std::vector<model*> models;
    std::vector<const clade*> order;        // reverse level order, the order inference_prune visits
    std::vector<const clade*> internal;     // internal nodes in that order (root last)

    explicit setup(const args_t& a)
    {
        in.tree_file_path = a.str("tree");
        in.input_file_path = a.str("fam");
        in.error_model_file_path = a.str("err");
        in.use_error_model = a.has("err") || a.integer("esterr", 0) != 0;      // --esterr 1: "-e" without a file => epsilon is estimated
        in.lambda_tree_file_path = a.str("ltree");
        in.rootdist = a.str("rootdist");
        in.n_gamma_cats = (int)a.integer("k", 1);
        in.fixed_alpha = a.num("alpha", -1.0);
        in.exclude_zero_root_families = a.integer("filter", 1) != 0;
        std::string lam = a.str("lambda");
        if (!lam.empty()) {
            if (lam.find(',') != std::string::npos || a.has("ltree")) in.fixed_multiple_lambdas = lam;
            else in.fixed_lambda = atof(lam.c_str());
        }
        if (a.has("poisson")) { in.use_uniform_eq_freq = false; in.poisson_lambda = a.num("poisson", 0.0); }

        data.read_datafiles(in);
        if (a.has("maxfam")) { data.max_family_size = (int)a.integer("maxfam", 0); }
        if (a.has("maxroot")) { data.max_root_family_size = (int)a.integer("maxroot", 0); }
        if (in.exclude_zero_root_families) {
            auto rem = std::remove_if(data.gene_families.begin(), data.gene_families.end(), [this](const gene_family& fam) {
                return !fam.exists_at_root(data.p_tree);
            });
            data.gene_families.erase(rem, data.gene_families.end());
        }
        long limit = a.integer("limit", -1);
        if (limit >= 0 && (size_t)limit < data.gene_families.size()) data.gene_families.resize(limit);
        data.p_prior.reset(root_eq_dist_factory(in, &data.gene_families));
#ifdef WITH_CUDA_MODELS
        models = a.integer("cuda", 0) ? build_cuda_models(in, data) : build_models(in, data);
#else
        if (a.integer("cuda", 0)) throw std::runtime_error("--cuda needs the ref_harness_cuda build");
        models = build_models(in, data);
#endif

        data.p_tree->apply_reverse_level_order([this](const clade* c) { order.push_back(c); if (!c->is_leaf()) internal.push_back(c); });
    }
};

void print_setup(const setup& s)
{
    printf("\"n_families\": %zu, \"max_family_size\": %d, \"max_root_family_size\": %d, \"threads\": %d, ",
        s.data.gene_families.size(), s.data.max_family_size, s.data.max_root_family_size, omp_get_max_threads());
    printf("\"node_order\": [");
    for (size_t i = 0; i < s.order.size(); ++i) printf("%s\"%s\"", i ? ", " : "", s.order[i]->get_taxon_name().c_str());
    printf("], ");
}

int cmd_bd(const args_t& a)
{
    double v = the_probability_of_going_from_parent_fam_size_to_c(a.num("lambda", 0), a.num("t", 0), (int)a.integer("s", 0), (int)a.integer("c", 0));
    printf("{\"p\": "); jnum(v); printf("}\n");
    return 0;
}

int cmd_bdlog(const args_t& a)
{
    double v = birthdeath_rate_with_log_alpha((int)a.integer("s", 0), (int)a.integer("c", 0), a.num("logalpha", 0), a.num("coeff", 0));
    printf("{\"p\": "); jnum(v); printf("}\n");
    return 0;
}

int cmd_matrix(const args_t& a)
{
    int n = (int)a.integer("n", 0);
    double lambda = a.num("lambda", 0), t = a.num("t", 0);
    matrix_cache cache(n);
    cache.precalculate_matrices({ lambda }, std::set<double>{ t });
    const matrix* m = cache.get_matrix(t, lambda);
    std::vector<double> flat((size_t)n * n);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) flat[(size_t)i * n + j] = m->get(i, j);
    dumper d(a.str("dump"));
    d.doubles(flat.data(), flat.size());
    matrix_cache_key key(n, lambda, t);
    printf("{\"n\": %d, \"lambda_q\": ", n); jnum(key.lambda()); printf(", \"t_q\": "); jnum(key.branch_length());
    printf(", \"saturated\": %d", matrix_cache::is_saturated(t, lambda) ? 1 : 0);
    if (n <= 16) { printf(", "); jarr("m", flat); }
    printf("}\n");
    return 0;
}

int cmd_gamma(const args_t& a)
{
    int k = (int)a.integer("k", 4);
    std::vector<double> freq(k), rate(k);
    get_gamma(freq, rate, a.num("alpha", 1.0));
    printf("{"); jarr("freq", freq); printf(", "); jarr("rate", rate); printf("}\n");
    return 0;
}

int cmd_poisson(const args_t& a)
{
    int n = (int)a.integer("n", 10);
    ::poisson_distribution pd(a.num("lambda", 1.0));
    root_distribution rd; rd.vectorize_uniform(n);
    pd.initialize(&rd);
    std::vector<double> v(n + 2);
    for (int i = 0; i < n + 2; ++i) v[i] = (double)pd.compute(i);
    printf("{"); jarr("prior", v); printf("}\n");
    return 0;
}

// inference_prune for the first --limit families with a lambda multiplier: dumps root vectors.
int cmd_prune(const args_t& a)
{
    setup s(a);
    model* m = s.models[0];
    double mult = a.num("mult", 1.0);
    matrix_cache calc(std::max(s.data.max_root_family_size, s.data.max_family_size) + 1);
    std::unique_ptr<lambda> ml(m->get_lambda()->multiply(mult));
    calc.precalculate_matrices(get_lambda_values(ml.get()), s.data.p_tree->get_branch_lengths());
    dumper d(a.str("dump"));
    printf("{"); print_setup(s);
    printf("\"root\": [");
    for (size_t i = 0; i < s.data.gene_families.size(); ++i) {
        auto v = inference_prune(s.data.gene_families[i], calc, m->get_lambda(), s.data.p_error_model, s.data.p_tree, mult,
            s.data.max_root_family_size, s.data.max_family_size);
        d.doubles(v.data(), v.size());
        if (s.data.gene_families.size() <= 8) {
            printf("%s[", i ? ", " : "");
            for (size_t j = 0; j < v.size(); ++j) { if (j) printf(", "); jnum(v[j]); }
            printf("]");
        }
    }
    printf("]}\n");
    return 0;
}

// compute_pvalues with a fixed seed.  The reference keeps the simulated families to itself, so after the call the
// same random stream is replayed through the same reference functions (set_weighted_random_family_size,
// compute_node_probability) to export what was simulated: leaf counts, unsorted conditional distributions, and the
// observed families' likelihoods.  "replay_matches" says the replay reproduced the reference's p-values exactly.
int cmd_pvalues(const args_t& a)
{
    setup s(a);
    model* m = s.models[0];
    const int nsim = (int)a.integer("nsim", 100);
    const unsigned seed = (unsigned)a.integer("seed", 10);
    const int mf = s.data.max_family_size, mrf = s.data.max_root_family_size;
    const clade* tree = s.data.p_tree;
    const lambda* lam = m->get_lambda();
    matrix_cache cache(std::max(mrf, mf) + 1);
    cache.precalculate_matrices(get_lambda_values(lam), tree->get_branch_lengths());

    randomizer_engine.seed(seed);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<double> pv;
#ifdef WITH_CUDA_MODELS
    if (a.integer("cuda", 0)) pv = compute_pvalues_cuda(tree, s.data.gene_families, lam, cache, nsim, mf, mrf);
    else
#endif
    pv = compute_pvalues(tree, s.data.gene_families, lam, cache, nsim, mf, mrf);
    double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    if (a.integer("replay", 1) == 0) {       // timing runs: skip the (serial, CPU) replay
        printf("{"); print_setup(s);
        printf("\"nsim\": %d, \"seed\": %u, \"seconds\": %.6f}\n", nsim, seed, seconds);
        return 0;
    }
    // ---- replay
    std::vector<const clade*> leaves;
    for (auto c : s.order) if (c->is_leaf()) leaves.push_back(c);
    std::vector<int> sim_counts((size_t)mrf * nsim * leaves.size());
    std::vector<double> cond((size_t)mrf * nsim);
    randomizer_engine.seed(seed);
    for (int root_size = 0; root_size < mrf; ++root_size) {
        std::vector<gene_family> fams(nsim);
        for (int i = 0; i < nsim; ++i) {
            clademap<int> sizes;
            sizes[tree] = root_size;
            auto fn = [&](const clade* c) { set_weighted_random_family_size(c, &sizes, lam, NULL, mf, cache); };
            tree->apply_prefix_order(fn);
            for (size_t l = 0; l < leaves.size(); ++l) {
                fams[i].set_species_size(leaves[l]->get_taxon_name(), sizes.at(leaves[l]));
                sim_counts[((size_t)root_size * nsim + i) * leaves.size() + l] = sizes.at(leaves[l]);
            }
        }
        for (int i = 0; i < nsim; ++i) {
            clademap<std::vector<double>> pruner;
            tree->apply_reverse_level_order([&](const clade* node) { pruner[node].resize(node->is_root() ? mrf : mf + 1); });
            tree->apply_reverse_level_order([&](const clade* c) { compute_node_probability(c, fams[i], NULL, pruner, mrf, mf, lam, cache); });
            cond[(size_t)root_size * nsim + i] = *std::max_element(pruner.at(tree).begin(), pruner.at(tree).end());
        }
    }
    std::vector<double> observed(s.data.gene_families.size());
    for (size_t f = 0; f < observed.size(); ++f) {
        clademap<std::vector<double>> pruner;
        tree->apply_reverse_level_order([&](const clade* node) { pruner[node].resize(node->is_root() ? mrf : mf + 1); });
        tree->apply_reverse_level_order([&](const clade* c) { compute_node_probability(c, s.data.gene_families[f], NULL, pruner, mrf, mf, lam, cache); });
        observed[f] = *std::max_element(pruner.at(tree).begin(), pruner.at(tree).end());
    }
    bool same = true;
    {
        std::vector<std::vector<double>> sorted(mrf);
        for (int r = 0; r < mrf; ++r) { sorted[r].assign(cond.begin() + (size_t)r * nsim, cond.begin() + (size_t)(r + 1) * nsim); std::sort(sorted[r].begin(), sorted[r].end()); }
        for (size_t f = 0; f < observed.size(); ++f) {
            double best = 0;
            for (int r = 0; r < mrf; ++r) best = std::max(best, pvalue(observed[f], sorted[r]));
            if (best != pv[f]) same = false;
        }
    }
    dumper d(a.str("dump"));
    d.ints(sim_counts.data(), sim_counts.size());
    d.doubles(cond.data(), cond.size());
    d.doubles(observed.data(), observed.size());
    d.doubles(pv.data(), pv.size());
    printf("{"); print_setup(s);
    printf("\"nsim\": %d, \"seed\": %u, \"n_leaves\": %zu, \"seconds\": %.6f, \"replay_matches\": %s, \"leaf_order\": [", nsim, seed, leaves.size(), seconds, same ? "true" : "false");
    for (size_t l = 0; l < leaves.size(); ++l) printf("%s\"%s\"", l ? ", " : "", leaves[l]->get_taxon_name().c_str());
    printf("]");
    if (pv.size() <= 16) { printf(", "); jarr("pvalues", pv); }
    printf("}\n");
    return 0;
}

// One (or --reps) model::infer_family_likelihoods; optional reconstruction.
int cmd_eval(const args_t& a)
{
    setup s(a);
    model* m = s.models[0];
    int reps = (int)a.integer("reps", 1);
    double score = 0, best = 1e300, total = 0;
    for (int r = 0; r < reps; ++r) {
        auto t0 = std::chrono::steady_clock::now();
        score = m->infer_family_likelihoods(s.data.p_prior.get(), s.data.rootdist, m->get_lambda());
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        best = std::min(best, dt); total += dt;
    }
#ifdef WITH_CUDA_MODELS
    // the CUDA models build model::results on demand (write_family_likelihoods); the dumps below read it directly
    if (auto cb = dynamic_cast<cuda_base_model*>(m)) cb->materialize_results();
    if (auto cg = dynamic_cast<cuda_gamma_model*>(m)) cg->materialize_results();
    if (a.integer("cuda", 0) && a.has("viterbi")) cuda_models_use_device_branch_probabilities(true);
#endif
    dumper d(a.str("dump"));
    size_t F = s.data.gene_families.size();
    auto gm = dynamic_cast<gamma_model*>(m);
    printf("{"); print_setup(s);
    printf("\"model\": \"%s\", \"score\": ", m->name().c_str()); jnum(score);
    printf(", \"seconds_best\": %.6f, \"seconds_mean\": %.6f, \"reps\": %d", best, total / reps, reps);
    if (gm) {
        printf(", "); jarr("multipliers", gm->_lambda_multipliers); printf(", "); jarr("cat_probs", gm->_gamma_cat_probs);
        // dump layout: F x k category likelihoods (as left in _category_likelihoods, possibly truncated on failure)
        size_t k = gm->_gamma_cat_probs.size();
        std::vector<double> flat(F * k, std::nan(""));
        const std::vector<std::vector<double>>* cl = &gm->_category_likelihoods;
#ifdef WITH_CUDA_MODELS
        if (auto cg = dynamic_cast<cuda_gamma_model*>(m)) cl = &cg->category_likelihoods();
#endif
        for (size_t i = 0; i < F && i < cl->size(); ++i) for (size_t j = 0; j < (*cl)[i].size() && j < k; ++j) flat[i * k + j] = (*cl)[i][j];
        d.doubles(flat.data(), flat.size());
    }
    else {
        std::vector<double> lnl(F);
        for (size_t i = 0; i < F && i < m->results.size(); ++i) lnl[i] = m->results[i].posterior_probability;
        d.doubles(lnl.data(), lnl.size());
    }
    if (a.has("recon")) {
        matrix_cache calc(std::max(s.data.max_root_family_size, s.data.max_family_size) + 1);
        auto t0 = std::chrono::steady_clock::now();
        std::unique_ptr<reconstruction> rec(m->reconstruct_ancestral_states(s.data.gene_families, &calc, s.data.p_prior.get()));
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf(", \"recon_seconds\": %.6f, \"internal_order\": [", dt);
        for (size_t i = 0; i < s.internal.size(); ++i) printf("%s\"%s\"", i ? ", " : "", s.internal[i]->get_taxon_name().c_str());
        printf("]");
        dumper dr(a.str("dumprecon"));
        std::vector<int> row;
        if (gm) {
            auto gr = dynamic_cast<gamma_model_reconstruction*>(rec.get());
            size_t k = gm->_gamma_cat_probs.size();
            for (auto& fam : s.data.gene_families) {
                auto& r = gr->_reconstructions.at(fam.id());
                row.clear();
                for (size_t c = 0; c < k; ++c) for (auto node : s.internal) row.push_back(r.category_reconstruction[c].at(node));
                dr.ints(row.data(), row.size());
            }
        }
        else {
            auto br = dynamic_cast<base_model_reconstruction*>(rec.get());
            for (auto& fam : s.data.gene_families) {
                auto& r = br->_reconstructions.at(fam.id());
                row.clear();
                for (auto node : s.internal) row.push_back(r.at(node));
                dr.ints(row.data(), row.size());
            }
        }
        if (a.has("viterbi")) {
            // compute_viterbi_sum for every family and node (src/execute.cpp:165-176 with every family selected):
            // dump = ints [F][nodes] reconstructed sizes (node_order), then doubles [F][nodes] probabilities (-1 = invalid)
            const int mf = s.data.max_family_size;
            std::vector<double> pv(F, 0.0);
            branch_probabilities probs;
#ifdef WITH_CUDA_MODELS
            if (a.integer("cuda", 0))
                probs = compute_branch_probabilities_cuda(s.data.p_tree, s.data.gene_families, rec.get(), pv, 1.0, m->get_lambda(), mf, s.data.max_root_family_size);
            else
#endif
            for (auto& fam : s.data.gene_families)
                for (auto c : s.order) probs.set(fam, c, compute_viterbi_sum(c, fam, rec.get(), mf, calc, m->get_lambda()));
            dumper dv(a.str("dumpviterbi"));
            std::vector<int> sz;
            std::vector<double> val;
            for (auto& fam : s.data.gene_families)
                for (auto c : s.order) {
                    sz.push_back(rec->reconstructed_size(fam, c));
                    auto bp = probs.at(fam, c);
                    val.push_back(bp._is_valid ? bp._value : -1.0);
                }
            dv.ints(sz.data(), sz.size());
            dv.doubles(val.data(), val.size());
        }
    }
    printf("}\n");
    return 0;
}

// optimizer::optimize over the scorer the model hands out (src/execute.cpp:78-104).
int cmd_fit(const args_t& a)
{
    randomizer_engine.seed((unsigned)a.integer("seed", 10));
    setup s(a);
    model* m = s.models[0];
    std::unique_ptr<inference_optimizer_scorer> scorer(m->get_lambda_optimizer(s.data));
    if (!scorer) throw std::runtime_error("nothing to optimise");
    struct counting : optimizer_scorer {
        inference_optimizer_scorer* inner; int evals = 0; double seconds = 0, first = 0;
        std::vector<double> initial_guesses() override { return inner->initial_guesses(); }
        double calculate_score(const double* v) override {
            auto t0 = std::chrono::steady_clock::now();
            double r = inner->calculate_score(v); ++evals;
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            seconds += dt;
            if (evals == 1) first = dt;         // includes one-time set-up (de-duplication; for the CUDA models context creation + upload)
            return r;
        }
    } cs; cs.inner = scorer.get();
    optimizer opt(&cs);
    opt.quiet = true;
    auto t0 = std::chrono::steady_clock::now();
    auto result = opt.optimize(s.in.optimizer_params);
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    scorer->finalize(&result.values[0]);
    printf("{"); print_setup(s);
    printf("\"model\": \"%s\", ", m->name().c_str()); jarr("values", result.values);
    printf(", \"score\": "); jnum(result.score);
    printf(", \"iterations\": %d, \"evaluations\": %d, \"seconds\": %.6f, \"seconds_in_score\": %.6f, \"first_evaluation_seconds\": %.6f", result.num_iterations,
           cs.evals, dt, cs.seconds, cs.first);
#ifdef WITH_CUDA_MODELS
    {
        cuda_bridge* br = nullptr;
        if (auto cb = dynamic_cast<cuda_base_model*>(m)) br = &cb->bridge();
        if (auto cg = dynamic_cast<cuda_gamma_model*>(m)) br = &cg->bridge();
        if (br) {
            const std::vector<double> hs = br->host_seconds();
            printf(", \"devices\": %d, \"device_evaluations\": %ld, \"device_seconds\": %.6f, \"bind_seconds\": %.6f, \"eval_call_seconds\": %.6f, "
                   "\"staging_seconds\": %.6f, \"enqueue_seconds\": %.6f, \"wait_seconds\": %.6f",
                   br->device_count(), br->evaluations(), br->device_seconds(), hs[0], hs[1], hs[2], hs[3], hs[4]);
        }
    }
#endif
    printf("}\n");
    return 0;
}

}

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: ref_harness bd|bdlog|matrix|gamma|poisson|prune|eval|fit --key value ...\n"); return 2; }
    init_lgamma_cache();
    try {
        std::string cmd = argv[1];
        args_t a = parse(argc, argv, 2);
        if (cmd == "bd") return cmd_bd(a);
        if (cmd == "bdlog") return cmd_bdlog(a);
        if (cmd == "matrix") return cmd_matrix(a);
        if (cmd == "gamma") return cmd_gamma(a);
        if (cmd == "poisson") return cmd_poisson(a);
        if (cmd == "prune") return cmd_prune(a);
        if (cmd == "eval") return cmd_eval(a);
        if (cmd == "pvalues") return cmd_pvalues(a);
        if (cmd == "fit") return cmd_fit(a);
        fprintf(stderr, "unknown command %s\n", cmd.c_str());
        return 2;
    }
    catch (std::exception& e) {
        fprintf(stderr, "ref_harness: %s\n", e.what());
        return 1;
    }
}
