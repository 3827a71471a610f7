/* TEST INFRASTRUCTURE — CPU restatement of the reference's per-family birth-death likelihood path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; nothing under cafexp_b200/ does.  Every function cites the reference
 * file:line (relative to /root/reference) whose behaviour it restates.  Parity is PINNED: the
 * restatement is checked against golden vectors produced by the compiled, unmodified reference
 * (oracle/_ref/ref_harness, scripts/make_golden.py -> tests/golden/) in tests/test_oracle_golden.py.
 *
 * Plain C11, doubles only, no FMA contraction (-ffp-contract=off), sums in the reference's order.
 */
#ifndef CAFE_ORACLE_H
#define CAFE_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Flattened species tree.  Nodes are numbered in the order clade::apply_reverse_level_order
 * (src/clade.cpp:255-280) visits them, so children always precede parents and the root is last. */
typedef struct orc_tree {
    int n_nodes;
    const int* parent;        /* [n_nodes], -1 for the root                                   */
    const int* child_offset;  /* [n_nodes+1] CSR into child_list, children in Newick order    */
    const int* child_list;    /* [n_nodes-1]                                                  */
    const int* leaf_col;      /* [n_nodes] column of the count matrix for leaves, -1 otherwise */
    const double* branch;     /* [n_nodes] raw branch length (root: unused)                   */
    const int* lambda_index;  /* [n_nodes] 0-based lambda index of the branch above the node  */
} orc_tree;

enum { ORC_BASE_LOGMAX = 0, ORC_GAMMA_LINSUM = 1 };

/* src/probability.cpp:66-88 */
void   orc_init(void);
double orc_chooseln(double n, double r);
/* src/probability.cpp:101-145 */
double orc_birthdeath_rate_with_log_alpha(int s, int c, double log_alpha, double coeff);
/* src/probability.cpp:147-164 */
double orc_bd_probability(double lambda, double branch_length, int parent_size, int size);
/* src/matrix_cache.h:47-60 */
double orc_quantise_lambda(double lambda);
double orc_quantise_branch(double t);
/* src/matrix_cache.cpp:115-119 */
int    orc_is_saturated(double branch_length, double lambda);
/* src/matrix_cache.cpp:70-77,121-171: one n x n row-major matrix from RAW (lambda, t) */
void   orc_build_matrix(int n, double lambda_raw, double t_raw, double* out);

/* src/gamma.cpp:15-241 (PAML discrete gamma, mean-of-category) */
double orc_incomplete_gamma(double x, double alpha, double ln_gamma_alpha);
double orc_point_normal(double prob);
double orc_point_chi2(double prob, double v);
void   orc_get_gamma(int k, double alpha, double* freq, double* rate);

/* src/root_distribution.cpp:15-31 + src/root_equilibrium_distribution.cpp:20-32.
 * rootdist_sizes/counts: the map<int,int> entries in key order, n_entries==0 -> vectorize_uniform(mrf).
 * Writes (double)(float) prior values for val = 0..n_out-1. */
void   orc_prior_uniform(const int* rootdist_sizes, const int* rootdist_counts, int n_entries, int max_root_family_size,
                         double* out, int n_out);
/* src/root_equilibrium_distribution.cpp:47-50, .h:45-51, src/poisson.cpp:19-36 */
void   orc_prior_poisson(double poisson_lambda, const int* rootdist_sizes, const int* rootdist_counts, int n_entries,
                         int max_root_family_size, double* out, int n_out);

/* src/error_model.cpp:79-109: rewrite a dense [rows][3] table for a new epsilon; returns 0 or -1 */
int    orc_error_model_replace_epsilon(double* probs, int rows, double old_eps, double new_eps);

/* src/core.cpp:133-144 + src/probability.cpp:173-242: one family, one lambda multiplier.
 * lambdas: [n_lambdas] RAW values already multiplied by the category multiplier.
 * err: dense [err_rows][err_ndev] deviation probabilities indexed by observed count, or NULL.
 * out_root: [mrf], index j <-> root size j+1.  Returns 0, or -1 on a bad count. */
int    orc_inference_prune(const orc_tree* tree, const int32_t* counts_row, const double* lambdas, int n_lambdas,
                           const double* err, int err_rows, int err_ndev, int max_family_size, int max_root_family_size,
                           double* out_root);

/* src/base_model.cpp:53-112 (mode ORC_BASE_LOGMAX) / src/gamma_core.cpp:144-248 (ORC_GAMMA_LINSUM).
 * lambdas: [k][n_lambdas] RAW, i.e. lambda_i * multiplier_k.  prior: [mrf] as (double)(float).
 * family_lnl: [F] or NULL; cat_lk: [F][k] or NULL (gamma only); failed: [F] flags or NULL.
 * Returns -lnL (+inf when the reference would). */
double orc_infer_family_likelihoods(const orc_tree* tree, const int32_t* counts, int64_t n_families, int n_leaves,
                                    const double* lambdas, int n_lambdas, const double* cat_probs, int k,
                                    const double* prior, const double* err, int err_rows, int err_ndev,
                                    int max_family_size, int max_root_family_size, int mode,
                                    double* family_lnl, double* cat_lk, uint8_t* failed, int64_t* n_failed);

/* src/gene_family_reconstructor.cpp:13-165 for every family and category.
 * prior: [>= min(mf,mrf)+1] as (double)(float), indexed by root SIZE (not size-1).
 * states: [F][k][n_internal], internal nodes in tree order (root last). */
int    orc_reconstruct(const orc_tree* tree, const int32_t* counts, int64_t n_families, int n_leaves,
                       const double* lambdas, int n_lambdas, int k, const double* prior,
                       int max_family_size, int max_root_family_size, int32_t* states);

/* src/gamma_core.cpp:282-299 */
void   orc_weighted_averages(const int32_t* states, int k, int n_internal, const double* cat_probs, double* out);

/* src/probability.cpp:301-308, :396-399: max of the root vector per family (one lambda set, no prior). */
int    orc_root_max(const orc_tree* tree, const int32_t* counts, int64_t n_families, int n_leaves, const double* lambdas, int n_lambdas,
                    int max_family_size, int max_root_family_size, double* out);
/* src/probability.cpp:379-389 */
double orc_pvalue(double v, const double* sorted, int n);
/* src/probability.cpp:310 + :391-409; cond [n_root_sizes][n_sim] is sorted in place */
void   orc_pvalues(double* cond, int n_root_sizes, int n_sim, const double* observed, int64_t n_families, double* pvalues);

/* src/gene_family_reconstructor.cpp:361-400 for every (family, node): node_sizes [F][n_nodes] reconstructed sizes in tree
 * order, selected [F] or NULL, out [F][n_nodes] with -1 where the reference has no value. */
void   orc_branch_probabilities(const orc_tree* tree, const int32_t* node_sizes, const uint8_t* selected, int64_t n_families,
                                const double* lambdas, int n_lambdas, int max_family_size, int max_root_family_size, double* out);

#ifdef __cplusplus
}
#endif
#endif
