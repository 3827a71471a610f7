/* cafe_b200.h — C ABI of the B200-native (sm_100a) per-family birth-death likelihood engine.
 *
 * This is the drop-in boundary for ONE hot path of Han9527/CAFExp (CAFE5): everything
 * model::infer_family_likelihoods and model::reconstruct_ancestral_states compute below the host
 * orchestration.  The reference has no FFI layer; its seam is the C++ virtuals
 *      model::infer_family_likelihoods          src/core.h:171
 *      model::reconstruct_ancestral_states      src/core.h:179
 *      optimizer_scorer::calculate_score        src/optimizer_scorer.h:22
 * and the entry points below are what CUDA-backed subclasses of base_model / gamma_model call
 * (integration/cuda_models.cpp; binding notes in INTEGRATION.md).
 *
 * Conventions
 *   - plain C, opaque handle, int status (0 = ok, <0 = error; cafe_b200_last_error() has the text),
 *     no exceptions cross the boundary, every output buffer is caller-owned HOST memory unless the
 *     name says _device.
 *   - one host thread per context (as the reference: every call comes from the Nelder-Mead thread).
 *   - numerical failure is reported in-band like the reference: *neg_lnl = +inf and *n_failed > 0,
 *     never as an error status.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns
 *     CAFE_B200_ERR_CUDA.
 */
#ifndef CAFE_B200_H
#define CAFE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAFE_B200_ABI_VERSION 2

enum {
    CAFE_B200_OK = 0,
    CAFE_B200_ERR_ARG = -1,         /* bad argument (null pointer, size, tree shape)                 */
    CAFE_B200_ERR_CUDA = -2,        /* CUDA runtime error or no device                               */
    CAFE_B200_ERR_COUNT_RANGE = -3, /* a leaf count (+ error-model deviation) is outside 0..max_family_size:
                                       the reference would index past its vectors (src/probability.cpp:191,197) */
    CAFE_B200_ERR_LIMIT = -4        /* size beyond what this build supports (see cafe_b200_limits)   */
};

/* mode of cafe_b200_eval */
enum {
    CAFE_B200_BASE_LOGMAX = 0,  /* base_model: lnL_i = max_j(log L_i[j] + log prior[j])         src/base_model.cpp:89-106 */
    CAFE_B200_GAMMA_LINSUM = 1, /* gamma_model: lnL_i = log sum_k max_j(L_ik[j] prior[j]) catprob[k], fail if sum_j L_ik[j]==0
                                                                                                  src/gamma_core.cpp:144-166,203-219 */
    CAFE_B200_ROOT_MAX = 2      /* internal to cafe_b200_root_max: lk_i = max_j L_i[j], no prior  src/probability.cpp:308,399 */
};

/* options for cafe_b200_set_option */
enum {
    CAFE_B200_OPT_RESCALE = 1,  /* 0 (default): reference arithmetic, partial likelihoods may underflow exactly as in the
                                   reference.  1: exact power-of-two per-family rescaling of every internal-node vector;
                                   identical results wherever the reference does not underflow, and the reference's
                                   "all zero" failure verdict is reproduced from the tracked exponent. */
    CAFE_B200_OPT_MAX_SLOTS = 2 /* cap the on-chip vector storage per thread block (>= 2; default: as much as fits): the
                                   reconstruction kernel keeps `value` shared-memory slots, the pruning kernel parks at most
                                   `value - 2` partial products in tensor memory.  Less forces spills to device scratch; used by tests. */
};

/* Species tree, flattened.  Node numbering = the order clade::apply_reverse_level_order visits
 * (src/clade.cpp:255-280): children before parents, root last.  Children in Newick order
 * (clade::_descendants).  n-ary nodes are allowed. */
typedef struct cafe_b200_tree {
    int n_nodes;
    const int* parent;         /* [n_nodes]  -1 for the root                                              */
    const int* child_offset;   /* [n_nodes+1] CSR offsets into child_list                                  */
    const int* child_list;     /* [n_nodes-1]                                                              */
    const int* leaf_col;       /* [n_nodes]  column of the count matrix for a leaf, -1 for internal nodes  */
    const double* branch;      /* [n_nodes]  RAW branch length; quantised inside as matrix_cache_key does
                                             (long(t*1000)/1000.0, src/matrix_cache.h:50,58)               */
    const int* lambda_index;   /* [n_nodes]  0-based index of the lambda that applies to the branch above
                                             the node (multiple_lambda's node-name map, src/lambda.cpp:32-40) */
} cafe_b200_tree;

typedef struct cafe_b200_ctx cafe_b200_ctx;

typedef struct cafe_b200_limits {
    int max_matrix_size;       /* max(max_family_size, max_root_family_size)+1 supported      */
    int max_categories;
    int max_nodes;
    int families_per_tile;     /* most families pruned together as matrix columns by one thread block */
} cafe_b200_limits;

int  cafe_b200_abi_version(void);
void cafe_b200_get_limits(cafe_b200_limits* out);

/* Number of CUDA devices visible (0 without a driver/GPU). */
int  cafe_b200_device_count(void);

/* Replaces the model constructor's view of (tree, families, max sizes): src/core.h:59-70.
 * leaf_counts: HOST [n_families][n_leaves] int32, column = tree->leaf_col; copied to the device.
 * device: CUDA ordinal.  The context owns one stream; all work of a call is enqueued on it. */
int  cafe_b200_create(cafe_b200_ctx** out, const cafe_b200_tree* tree, const int32_t* leaf_counts,
                      int64_t n_families, int n_leaves, int max_family_size, int max_root_family_size, int device);

/* The same over SEVERAL devices of one box, driven by the one host thread the reference's optimizer runs on
 * ("const int* devices, int n_devices"; section 8e).  Families are split into n_devices contiguous
 * ranges, device i owning [F*i/n, F*(i+1)/n); every other entry point then works on all shards at once: the host inputs
 * are staged once, each device builds its matrices and prunes its range concurrently, per-family outputs land in the
 * caller's arrays at the families' global positions, and the score is the sum of the devices' [sum lnL, #failed]
 * pairs, added on the host in device-list order (deterministic; 16 bytes per device, no collective library needed).
 *   leaf_counts  HOST [n_families][n_leaves], elements of count_bytes = 1 (uint8), 2 (uint16) or 4 (int32) bytes.
 *                On the device counts are 1 byte each when max_family_size <= 255 (else 2): handing over uint8
 *                moves a quarter of the bytes of int32 (pinned host memory makes the copy asynchronous). */
int  cafe_b200_create_multi(cafe_b200_ctx** out, const cafe_b200_tree* tree, const void* leaf_counts, int count_bytes,
                            int64_t n_families, int n_leaves, int max_family_size, int max_root_family_size,
                            const int* devices, int n_devices);
int  cafe_b200_n_devices(const cafe_b200_ctx* ctx);

/* Page-locked host memory for inputs / outputs of the calls below (plain malloc'ed memory works too, but is copied
 * through a driver staging buffer).  NULL when no device / out of memory. */
void* cafe_b200_alloc_pinned(size_t bytes);
void  cafe_b200_free_pinned(void* p);
void cafe_b200_destroy(cafe_b200_ctx* ctx);
const char* cafe_b200_last_error(const cafe_b200_ctx* ctx);   /* ctx may be NULL: last create() error */

/* Re-upload the count matrix (same shape as at create).  Replaces model::set_families (src/core.h:148).
 * The range check (0 <= count <= max_family_size) runs on the device after the copy: on CAFE_B200_ERR_COUNT_RANGE
 * the context holds the rejected matrix and refuses every evaluation until a valid one is set. */
int  cafe_b200_set_families(cafe_b200_ctx* ctx, const int32_t* leaf_counts, int64_t n_families);
int  cafe_b200_set_families_ex(cafe_b200_ctx* ctx, const void* leaf_counts, int count_bytes, int64_t n_families);

/* Leaf error model: dense HOST [rows][n_deviations] table indexed by OBSERVED count, i.e. row s =
 * error_model::get_probs(s) (src/error_model.cpp:52-57); deviations are centred, -(nd-1)/2..+(nd-1)/2
 * (src/probability.cpp:185).  probs == NULL removes the error model.  Cheap to call before every evaluation (the
 * epsilon optimiser edits the model in place): the table is uploaded only when it differs from the last one. */
int  cafe_b200_set_error_model(cafe_b200_ctx* ctx, const double* probs, int rows, int n_deviations);

int  cafe_b200_set_option(cafe_b200_ctx* ctx, int option, int value);

/* Distributed matrix build across PROCESSES (one single-device context per process / rank).  By
 * default every context builds every transition matrix of an evaluation itself; with a partition announced here it
 * builds only slab `part` of `n_parts` (keys are padded to a whole number per part) and then calls `gather`, which must
 * all-gather the slabs in place on the given stream — `matrices` is this context's matrix buffer (device memory), slab i
 * lives at byte offset i * slab_bytes, this rank's own slab is already in place (e.g. ncclAllGather / torch
 * all_gather_into_tensor with sendbuff = recvbuff + part * slab_bytes).  Every rank must evaluate with identical
 * parameters.  Returns 0 on success.  (A multi-device context does the same by itself with peer copies over NVLink.) */
typedef int (*cafe_b200_gather_fn)(void* user, void* matrices, size_t slab_bytes, int n_parts, void* cuda_stream);
int  cafe_b200_set_build_partition(cafe_b200_ctx* ctx, int part, int n_parts, cafe_b200_gather_fn gather, void* user);

/* Enqueue on an existing CUDA stream (cudaStream_t passed as void*; NULL = the legacy default stream)
 * instead of the context's own non-blocking stream.  Single-device contexts only. */
int  cafe_b200_set_stream(cafe_b200_ctx* ctx, void* cuda_stream);

/* One evaluation of the likelihood = the body of base_model::infer_family_likelihoods
 * (src/base_model.cpp:77-107) or gamma_model::infer_family_likelihoods (src/gamma_core.cpp:196-244):
 * builds every transition matrix for (branch, lambda, category) on the device, prunes every family,
 * applies root prior / category weights, reduces.
 *   lambdas    HOST [n_categories][n_lambdas] RAW values lambda_i * multiplier_k (multiply first as
 *              lambda::multiply does, src/lambda.h:47,80); quantised inside (long(x*1e9)/1e9).
 *   cat_probs  HOST [n_categories] (gamma_cat_probs; {1.0} for the base model)
 *   prior      HOST [max_root_family_size]: (double)prior->compute(j), j = 0..mrf-1 (a float widened)
 *   neg_lnl    out: -sum_i lnL_i, or +inf when the reference returns -log(0)
 *   family_lnl out HOST [n_families] or NULL (NaN for failed families)
 *   cat_lk     out HOST [n_families][n_categories] or NULL (gamma mode: _category_likelihoods)
 *   n_failed   out: families whose pruning "saturated" (gamma mode), may be NULL
 *   failed_idx out HOST [failed_cap] first indices of failed families, may be NULL */
int  cafe_b200_eval(cafe_b200_ctx* ctx, const double* lambdas, int n_lambdas, const double* cat_probs, int n_categories,
                    const double* prior, int mode, double* neg_lnl, double* family_lnl, double* cat_lk,
                    int64_t* n_failed, int64_t* failed_idx, int64_t failed_cap);

/* The category likelihoods [n_families][k] of the LAST cafe_b200_eval in gamma mode, for callers that passed
 * cat_lk = NULL there (an optimizer needs only the score; _category_likelihoods is read once, after the fit). */
int  cafe_b200_fetch_category_likelihoods(cafe_b200_ctx* ctx, int n_categories, double* cat_lk);

/* Same evaluation, asynchronous, result left on the device: result_device[0] = sum_i lnL_i over
 * non-failed families, result_device[1] = number of failed families (as a double).  This is the pair a
 * multi-process caller sum-allreduces (NCCL) across its family shards; no host synchronisation.  Single-device
 * contexts only (a multi-device context reduces inside cafe_b200_eval). */
int  cafe_b200_eval_device(cafe_b200_ctx* ctx, const double* lambdas, int n_lambdas, const double* cat_probs, int n_categories,
                           const double* prior, int mode, double* result_device);

/* Pupko joint ancestral reconstruction = reconstruct_gene_family for every family and category
 * (src/gene_family_reconstructor.cpp:13-165; callers src/base_model.cpp:145-162, src/gamma_core.cpp:301-347).
 *   prior_by_size HOST [min(mf,mrf)+1]: (double)prior->compute(size) — indexed by the size itself
 *   states  out HOST [n_families][n_categories][n_internal] int32, internal nodes in tree order (root last) */
int  cafe_b200_reconstruct(cafe_b200_ctx* ctx, const double* lambdas, int n_lambdas, int n_categories,
                           const double* prior_by_size, int32_t* states);

/* Branch probabilities of the reconstructed size changes = compute_viterbi_sum for every (family, node)
 * (src/gene_family_reconstructor.cpp:361-400; caller src/execute.cpp:165-176), from the device-resident matrices.
 *   lambdas    HOST [n_lambdas] RAW (one set, no categories: the caller passes p_model->get_lambda())
 *   node_sizes HOST [n_families][n_nodes] int32: reconstruction::reconstructed_size(family, node) for every node in tree
 *              order (leaves: the observed count)
 *   selected   HOST [n_families] flags or NULL (= all): the reference only asks for families with pvalue < threshold
 *   out        HOST [n_families][n_nodes]: the probability, or -1 where the reference has no value
 *              (root, size equal to the parent's, family not selected) */
int  cafe_b200_branch_probabilities(cafe_b200_ctx* ctx, const double* lambdas, int n_lambdas, const int32_t* node_sizes,
                                    const uint8_t* selected, double* out);

/* ---- family p-values (compute_pvalues, src/probability.cpp:411-444) ------------------------- */

/* The statistic both halves of compute_pvalues use: lk_i = max_j of the root partial-likelihood vector of family i
 * under one lambda set (no rate categories, no prior) = *max_element(inference root vector), src/probability.cpp:308
 * (simulated families, get_random_probabilities) and :399 (observed families, compute_tree_pvalue).
 *   lambdas HOST [n_lambdas] RAW;  out HOST [n_families].
 * The caller simulates the families on the host with the reference's own generator (std::mt19937 randomizer_engine +
 * set_weighted_random_family_size, src/probability.cpp:320-351) so that the random stream stays the reference's, builds
 * a context over them and calls this; the observed families go through their own context. */
int  cafe_b200_root_max(cafe_b200_ctx* ctx, const double* lambdas, int n_lambdas, double* out);

/* p-value of every observed family against the simulated conditional distributions:
 *   cond      HOST [n_root_sizes][n_sim] likelihoods of the simulated families, row s = root size s, UNSORTED
 *             (sorted on the device, src/probability.cpp:310)
 *   observed  HOST [n_families] likelihoods of the observed families
 *   pvalues   out HOST [n_families] = max_s idx_s / n_sim, idx_s = upper_bound(cond[s], observed_i) - begin, or
 *             n_sim - 1 when no simulated value is greater (pvalue + compute_tree_pvalue, src/probability.cpp:379-409)
 * Stateless; n_sim <= 4096. */
int  cafe_b200_pvalues(int device, const double* cond, int n_root_sizes, int n_sim, const double* observed,
                       int64_t n_families, double* pvalues);

/* ---- inspection entry points used by the parity tests ------------------------------------- */

/* Transition matrices as the device built them: HOST out [n_categories][n_nodes][N][max_family_size+1]
 * (row = parent size, col = child size <= max_family_size; the root's block is zero), N = matrix size.
 * Replaces matrix_cache::precalculate_matrices + get_matrix (src/matrix_cache.cpp:121-171, :80-97). */
int  cafe_b200_build_matrices(cafe_b200_ctx* ctx, const double* lambdas, int n_lambdas, int n_categories, double* out);
int  cafe_b200_matrix_size(const cafe_b200_ctx* ctx);

/* Root partial-likelihood vectors = inference_prune (src/core.cpp:133-144):
 * HOST out [n_families][n_categories][max_root_family_size], index j <-> root size j+1. */
int  cafe_b200_prune_roots(cafe_b200_ctx* ctx, const double* lambdas, int n_lambdas, int n_categories, double* out);

/* Host-only: flat reader of the CAFE tab-format family table ("Desc<TAB>Family ID<TAB>species..." then one line per family)
 * = read_gene_families (src/io.cpp:134-215) + gene_family::get_species_size, straight into the row-major count matrix the
 * engine takes.  Column l of the output is species leaf_names[l] (matched case-insensitively,
 * src/gene_family.h:10-25); other species columns are ignored; every leaf must have a column.
 *   counts  out HOST [cap_families][n_leaves] int32, or NULL to only count the families
 *   ids     out HOST [cap_families][id_stride] NUL-terminated family ids, or NULL
 * *n_families = families in the file (may exceed cap_families: call again with a larger buffer). */
int  cafe_b200_read_family_table(const char* path, const char* const* leaf_names, int n_leaves, int32_t* counts, int64_t cap_families,
                                 int64_t* n_families, char* ids, int id_stride);

/* Host-only (no GPU needed): the op list the reconstruction kernel walks for this tree with n_slots shared-memory
 * slots.  ops_out: [cap][4] = {type, a, b, node}; types: 0 LEAF_SET(a,node) 1 LEAF_MUL
 * 2 GEMM_SET(a,node) 3 GEMM_MUL(a,b,node) 4 SPILL(a->scratch b) 5 FILL(a<-scratch b) 6 (end of node) 7 ROOT(a). */
int  cafe_b200_plan_schedule(const cafe_b200_tree* tree, int n_slots, int* ops_out, int cap, int* n_ops, int* n_spill);

/* Host-only: the stack-machine program the pruning kernel walks.  ops_out: [cap][7] = {type, node, flags, stack,
 * leaf_begin, n_pre, n_post}; types: 0 LEAVES (vector = product of the n_pre leaf columns listed from leaf_begin),
 * 1 GEMM (acc = M(node) * vector; acc = [product of n_pre leaves *] [parked(stack) *, flags & 1] acc [* n_post leaves];
 * flags & 2: park acc at `stack`, else acc is the parent's vector), 2 ROOT.  leaves_out: leaf node ids. */
int  cafe_b200_plan_program(const cafe_b200_tree* tree, int* ops_out, int cap, int* n_ops, int* leaves_out, int leaves_cap,
                            int* n_leaf_refs, int* depth);

/* Host wall time (seconds) spent inside the library since create: [0] staging the evaluation parameters (key
 * quantisation, pow rows from libm, the program), [1] enqueueing copies and kernels, [2] waiting for the devices. */
int  cafe_b200_host_seconds(const cafe_b200_ctx* ctx, double* s3);

/* Human-readable description of the launch geometry chosen for this context (groups, ring, tensor-memory use). */
int  cafe_b200_describe(const cafe_b200_ctx* ctx, char* out, int cap);

/* Counters since create: kernel launches issued by this library, and evaluations. */
int64_t cafe_b200_launch_count(const cafe_b200_ctx* ctx);

/* Device time (ms, CUDA events on the context's streams, max over devices) of the phases of the LAST cafe_b200_eval /
 * cafe_b200_prune_roots / cafe_b200_reconstruct call: [0] matrix build, [1] pruning, [2] reduce,
 * [3] reconstruction.  Valid after the call returned (it synchronises). */
int  cafe_b200_last_timings(const cafe_b200_ctx* ctx, double* ms4);

/* The same for the most recent n calls, newest first: ms [n][4] (a caller that enqueues several cafe_b200_eval_device
 * calls without synchronising reads their kernel durations afterwards).  Returns the number of rows filled (<= 64);
 * call after the work has completed. */
int  cafe_b200_timing_history(const cafe_b200_ctx* ctx, int n, double* ms);

#ifdef __cplusplus
}
#endif
#endif
