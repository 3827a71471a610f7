// Reference-side binding of libcafe_b200.so: CUDA-backed subclasses of the reference's own models.
//
// This file is compiled INSIDE the reference's source tree (it includes the reference's headers; it holds
// none of the reference's code).  It is what a CAFExp maintainer adds to route the likelihood hot path to
// the B200 engine:
//
//      base_model  -> cuda_base_model      overrides infer_family_likelihoods      (src/base_model.cpp:53-112)
//                                                    reconstruct_ancestral_states  (src/base_model.cpp:145-162)
//      gamma_model -> cuda_gamma_model     overrides infer_family_likelihoods      (src/gamma_core.cpp:169-248)
//                                                    reconstruct_ancestral_states  (src/gamma_core.cpp:301-347)
//      build_models -> build_cuda_models   the single construction site            (src/core.cpp:16-50)
//
// Everything else — optimizer, scorers, lambda containers, error-model bookkeeping, discrete-gamma
// multipliers, report writers — is the reference's unchanged host code and keeps calling the same
// virtuals (optimizer_scorer::calculate_score -> model::infer_family_likelihoods, src/optimizer_scorer.cpp:19-33).
// There is no CPU fallback: if the CUDA library reports an error the overrides throw std::runtime_error,
// which cafexp() already catches (src/cafexp.cpp:215-218).
#ifndef CAFE_B200_CUDA_MODELS_H
#define CAFE_B200_CUDA_MODELS_H

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "base_model.h"
// src/gamma_core.h has no include guard; a translation unit that already included it defines this macro first
#ifndef CAFEXP_GAMMA_CORE_H_INCLUDED
#define CAFEXP_GAMMA_CORE_H_INCLUDED
#include "gamma_core.h"
#endif

struct cafe_b200_ctx;
class clade;
class gene_family;
class lambda;
class error_model;
class root_equilibrium_distribution;

//! Flattened (tree, de-duplicated families) and the device context that holds them.
//! One bridge per model object; it is rebuilt when the family vector it was built for changes.
//!
//! Devices: the CUDA ordinals listed in the environment variable CAFE_B200_DEVICES ("0,1,2,3", or "all"); default
//! device 0 (CAFE_B200_DEVICE is still honoured for a single ordinal).  With several devices the unique families are
//! split into contiguous ranges, one per device (cafe_b200_create_multi); nothing else changes for the caller.
class cuda_bridge {
public:
    cuda_bridge(const clade* p_tree, int max_family_size, int max_root_family_size);
    ~cuda_bridge();
    cuda_bridge(const cuda_bridge&) = delete;
    cuda_bridge& operator=(const cuda_bridge&) = delete;

    //! (Re)creates the device context if `families` is not the vector the current one was built from (same address,
    //! same size, same counts in a sample of its rows).  Call invalidate() after editing a bound vector in place.
    void bind(const std::vector<gene_family>& families);
    void invalidate() { _bound = nullptr; }

    //! Same for raw count rows [n_rows][leaves in leaf_nodes() order] (simulated families never become gene_family objects).
    void bind_rows(const std::vector<int>& rows, size_t n_rows);

    //! Hands the error model over as a dense [observed count][deviation] table (or removes it); the library uploads
    //! it only when it differs from the previous one, so this is called before every evaluation.
    void set_error_model(const error_model* p_error_model);

    //! lambdas[k][node] = (p_lambda * multiplier_k)->get_value_for_clade(node)   (src/lambda.h:45-48,76-84)
    std::vector<double> lambda_table(const lambda* p_lambda, const std::vector<double>& multipliers) const;

    //! prior[j] = (double)prior->compute(j), j = 0..n-1   (a float widened, src/root_equilibrium_distribution.h:15)
    static std::vector<double> prior_table(const root_equilibrium_distribution* prior, int n);

    //! One evaluation on the device.  family_lnl() then holds lnL per UNIQUE family (expand with unique_of()); the
    //! category likelihoods stay on the device until category_likelihoods() asks for them.
    //! Returns the number of unique families whose pruning failed (gamma mode); failed[] flags them.
    long evaluate(const std::vector<double>& lambdas, const std::vector<double>& cat_probs, const std::vector<double>& prior, int mode,
                  std::vector<char>& failed);
    const double* family_lnl() const { return _h_family_lnl; }
    //! [unique family][category] of the LAST gamma evaluation (fetched from the device on first use).
    const double* category_likelihoods();

    //! max_j of the root vector of every UNIQUE family under one lambda set (no categories, no prior): the
    //! likelihood compute_pvalues uses (src/probability.cpp:308, 399).
    std::vector<double> root_max(const std::vector<double>& lambdas);

    //! compute_viterbi_sum for every (bound row, node); node_sizes [rows][nodes in node order], selected [rows] or empty.
    std::vector<double> branch_probabilities(const std::vector<double>& lambdas, const std::vector<int>& node_sizes, const std::vector<unsigned char>& selected);

    //! Pupko reconstruction on the device: states[unique family][category][internal node].
    void reconstruct(const std::vector<double>& lambdas, int n_categories, const std::vector<double>& prior_by_size, std::vector<int>& states);

    size_t unique_of(size_t family_index) const { return _unique_of[family_index]; }
    size_t unique_count() const { return _n_unique; }
    int node_count() const { return (int)_order.size(); }
    const std::vector<const clade*>& internal_nodes() const { return _internal; }
    const std::vector<const clade*>& leaf_nodes() const { return _leaves; }       // = count-matrix columns
    const std::vector<const clade*>& nodes() const { return _order; }             // apply_reverse_level_order, root last
    int max_family_size() const { return _mf; }
    int max_root_family_size() const { return _mrf; }
    int device_count() const;
    //! device time (CUDA events: matrix build + pruning + reduction) and number of evaluate() calls so far
    //! incremented whenever the device context (and with it the last evaluation's outputs) is replaced
    unsigned long generation() const { return _generation; }
    double device_seconds() const { return _device_seconds; }
    long evaluations() const { return _evaluations; }
    //! host wall time so far: [0] bind (flatten + de-duplicate + create the context), [1] inside cafe_b200_eval, of which
    //! [2] staging, [3] enqueueing, [4] waiting for the devices
    std::vector<double> host_seconds() const;

    //! CUDA ordinals from CAFE_B200_DEVICES / CAFE_B200_DEVICE (see above).
    static std::vector<int> devices_from_environment();

private:
    const clade* _p_tree;
    int _mf, _mrf;
    std::vector<const clade*> _order;       // apply_reverse_level_order (src/clade.cpp:255-280)
    std::vector<const clade*> _internal;    // internal nodes in that order, root last
    std::vector<const clade*> _leaves;      // leaf nodes in that order = count-matrix columns
    std::vector<int> _parent, _child_offset, _child_list, _leaf_col, _lambda_index;
    std::vector<double> _branch;
    const std::vector<gene_family>* _bound = nullptr;
    size_t _bound_size = 0;
    size_t _bound_fingerprint = 0;
    std::vector<size_t> _unique_of;
    size_t _n_unique = 0;
    int _max_count = 0;
    cafe_b200_ctx* _ctx = nullptr;
    // page-locked result buffers, sized at bind
    double* _h_family_lnl = nullptr;
    double* _h_cat_lk = nullptr;
    size_t _cat_lk_capacity = 0;
    int _last_k = 0;
    bool _cat_lk_fetched = false;
    std::vector<long long> _failed_idx;
    unsigned long _generation = 0;
    double _device_seconds = 0.0;
    long _evaluations = 0;
    double _bind_seconds = 0.0, _eval_call_seconds = 0.0;
    double _lib_seconds_before[3] = {0.0, 0.0, 0.0};     // library counters of contexts already destroyed

    void check(int rc, const char* what) const;
    size_t fingerprint(const std::vector<gene_family>& families) const;
    void release();
};

//! Tell the drop-in that branch probabilities come from compute_branch_probabilities_cuda (below), i.e. that nobody
//! reads the host matrix_cache after reconstruct_ancestral_states (src/execute.cpp:158-170): the reconstruction then
//! skips the host-side precalculate_matrices (seconds of CPU work for k = 4).  Default: false (the cache is filled, as
//! an unmodified execute.cpp expects).
void cuda_models_use_device_branch_probabilities(bool on);

class cuda_base_model : public base_model {
public:
    cuda_base_model(lambda* p_lambda, const clade* p_tree, const std::vector<gene_family>* p_gene_families, int max_family_size,
                    int max_root_family_size, error_model* p_error_model);

    double infer_family_likelihoods(root_equilibrium_distribution* prior, const std::map<int, int>& root_distribution_map, const lambda* p_lambda) override;
    reconstruction* reconstruct_ancestral_states(const std::vector<gene_family>& families, matrix_cache* p_calc, root_equilibrium_distribution* p_prior) override;
    void write_family_likelihoods(std::ostream& ost) override;
    std::string name() const override { return "Base"; }

    //! Fills model::results from the last evaluation.  The reference rebuilds this vector in every evaluation
    //! (src/base_model.cpp:105) although only write_family_likelihoods reads it; here it is built on demand.
    void materialize_results();
    cuda_bridge& bridge() { return _bridge; }

private:
    cuda_bridge _bridge;
    bool _results_stale = false;
    unsigned long _generation_of_results = 0;      // bridge generation the pending results belong to
};

class cuda_gamma_model : public gamma_model {
public:
    cuda_gamma_model(lambda* p_lambda, clade* p_tree, std::vector<gene_family>* p_gene_families, int max_family_size, int max_root_family_size,
                     int n_gamma_cats, double fixed_alpha, error_model* p_error_model);
    cuda_gamma_model(lambda* p_lambda, clade* p_tree, std::vector<gene_family>* p_gene_families, int max_family_size, int max_root_family_size,
                     std::vector<double> gamma_categories, std::vector<double> multipliers, error_model* p_error_model);

    double infer_family_likelihoods(root_equilibrium_distribution* prior, const std::map<int, int>& root_distribution_map, const lambda* p_lambda) override;
    reconstruction* reconstruct_ancestral_states(const std::vector<gene_family>& families, matrix_cache* p_calc, root_equilibrium_distribution* p_prior) override;
    void write_family_likelihoods(std::ostream& ost) override;
    std::string name() const override { return "Gamma"; }

    //! category likelihoods of the last evaluation, per family (the reference keeps these private); built on demand
    const std::vector<std::vector<double>>& category_likelihoods();
    //! Fills model::results (F x k stashes, src/gamma_core.cpp:209-241) from the last evaluation, on demand.
    void materialize_results();
    cuda_bridge& bridge() { return _bridge; }

private:
    cuda_bridge _bridge;
    bool _explicit_categories;
    std::vector<double> _explicit_cat_probs;
    std::vector<std::vector<double>> _cat_lk;
    std::vector<double> _last_multipliers, _last_probs;
    bool _results_stale = false, _cat_lk_stale = false, _last_failed = false;
    unsigned long _generation_of_results = 0;

    std::vector<double> cat_probs() const;
};

//! Drop-in for compute_pvalues (src/probability.cpp:411-444; caller src/execute.cpp:161).  The families are
//! simulated on the host by the reference's own generator, in the reference's order (so the random stream, and
//! therefore every simulated family, is the reference's); their likelihoods and the observed families'
//! likelihoods come from the pruning kernel, the conditional distributions are sorted and searched on the device.
class matrix_cache;
std::vector<double> compute_pvalues_cuda(const clade* p_tree, const std::vector<gene_family>& families, const lambda* p_lambda,
                                         const matrix_cache& cache, int number_of_simulations, int max_family_size, int max_root_family_size);

//! Drop-in for the loop around compute_viterbi_sum (src/execute.cpp:165-176): branch probabilities of every family
//! whose p-value is below `test_pvalue`, for every node, from the device-resident transition matrices
//! (cafe_b200_branch_probabilities) instead of a host matrix_cache.
branch_probabilities compute_branch_probabilities_cuda(const clade* p_tree, const std::vector<gene_family>& families, const reconstruction* rec,
                                                       const std::vector<double>& pvalues, double test_pvalue, const lambda* p_lambda,
                                                       int max_family_size, int max_root_family_size);

//! Same decisions as build_models (src/core.cpp:16-50), instantiating the CUDA-backed subclasses.
std::vector<model*> build_cuda_models(const input_parameters& user_input, user_data& user_data);

#endif
