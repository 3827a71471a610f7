// C ABI (include/cafe_b200.h) and host orchestration of the sm_100a likelihood engine.
//
// Host work per evaluation is what the reference also does on the host before its hot loops:
// quantise (lambda, t) into matrix_cache keys (src/matrix_cache.h:47-60), derive alpha / coeff /
// log(alpha) (src/probability.cpp:150-156) — here once per unique key instead of once per matrix
// entry — and hand the device a flat schedule of the tree.  Everything O(families) or O(N^2) runs
// on the GPU.  There is no CPU fallback.
#include "../../include/cafe_b200.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "bd_matrix.cuh"
#include "common.cuh"
#include "prune.cuh"
#include "pupko.cuh"
#include "pvalue.cuh"
#include "reduce.cuh"
#include "viterbi.cuh"

using namespace cafe;

namespace {

std::string g_create_error;

struct Schedule {
    std::vector<Op> ops;
    int n_spill = 0;
};

struct FusedOp {
    int type, a, b, node, node2;
};

struct HostTree {
    int n_nodes = 0;
    std::vector<int> parent, child_offset, child_list, leaf_col, lambda_index;
    std::vector<double> branch;
    std::vector<long> branch_key;       // long(t * 1000)                         src/matrix_cache.h:50
    int n_internal = 0;
    int n_lambdas = 1;
    bool is_leaf(int v) const { return child_offset[v] == child_offset[v + 1]; }
};

// Post-order schedule with Sethi-Ullman ordering of internal children, so that the number of
// partial-likelihood vectors alive at once is the tree's Strahler-like "need"; vectors beyond the
// shared-memory slots are spilled to an L2-resident scratch area (rare: need <= log2(leaves)+1).
class ScheduleBuilder {
public:
    ScheduleBuilder(const HostTree& t, int slots) : tree(t), n_slots(slots), owner(slots, -1) { compute_need(); }

    Schedule build()
    {
        int root = tree.n_nodes - 1;
        int vid = emit(root);
        make_resident(vid, -1);
        out.ops.push_back({OP_ROOT, where[vid], 0, root});
        return out;
    }

private:
    const HostTree& tree;
    int n_slots;
    std::vector<int> owner;                 // physical slot -> vector id
    std::vector<int> where;                 // vector id -> slot (>= 0) or -(spill index + 1)
    std::vector<int> birth;                 // vector id -> creation order (victim choice: oldest)
    std::vector<int> free_spill;
    std::vector<int> need;
    Schedule out;
    int clock = 0;

    void compute_need()
    {
        need.assign(tree.n_nodes, 0);
        for (int v = 0; v < tree.n_nodes; ++v) {
            if (tree.is_leaf(v)) continue;
            std::vector<int> ns;
            for (int e = tree.child_offset[v]; e < tree.child_offset[v + 1]; ++e) {
                int c = tree.child_list[e];
                if (!tree.is_leaf(c)) ns.push_back(need[c]);
            }
            std::sort(ns.rbegin(), ns.rend());
            int n = 1;
            for (size_t i = 0; i < ns.size(); ++i) n = std::max(n, ns[i] + (i > 0 ? 1 : 0));
            need[v] = n;
        }
    }

    int new_vector()
    {
        where.push_back(-1000000);
        birth.push_back(clock++);
        return (int)where.size() - 1;
    }

    int acquire(int pin_a, int pin_b)
    {
        for (int s = 0; s < n_slots; ++s)
            if (owner[s] < 0) return s;
        int victim = -1;
        for (int s = 0; s < n_slots; ++s) {
            int vid = owner[s];
            if (vid == pin_a || vid == pin_b) continue;
            if (victim < 0 || birth[vid] < birth[owner[victim]]) victim = s;
        }
        int idx;
        if (!free_spill.empty()) { idx = free_spill.back(); free_spill.pop_back(); }
        else idx = out.n_spill++;
        out.ops.push_back({OP_SPILL, victim, idx, 0});
        where[owner[victim]] = -(idx + 1);
        owner[victim] = -1;
        return victim;
    }

    void make_resident(int vid, int pin)
    {
        if (where[vid] >= 0) return;
        int idx = -where[vid] - 1;
        int s = acquire(vid, pin);
        out.ops.push_back({OP_FILL, s, idx, 0});
        free_spill.push_back(idx);
        where[vid] = s;
        owner[s] = vid;
    }

    void release(int vid)
    {
        owner[where[vid]] = -1;
        where[vid] = -1000000;
    }

    // Binary nodes: internal child with the larger need first, leaves last (a*b == b*a exactly, so
    // the product is bit-identical to the reference's child order).  Nodes with more than two
    // children keep Newick order, because the reference multiplies factors in that order
    // (src/probability.cpp:211-217, src/gene_family_reconstructor.cpp:96-101) and a reassociated
    // product could differ in the last bit.
    int emit(int v)
    {
        std::vector<int> order;
        for (int e = tree.child_offset[v]; e < tree.child_offset[v + 1]; ++e) order.push_back(tree.child_list[e]);
        if (order.size() <= 2)
            std::stable_sort(order.begin(), order.end(), [this](int a, int b) {
                const int na = tree.is_leaf(a) ? -1 : need[a], nb = tree.is_leaf(b) ? -1 : need[b];
                return na > nb;
            });
        int acc = -1;
        for (int c : order) {
            if (!tree.is_leaf(c)) {
                int vid = emit(c);
                make_resident(vid, acc);
                if (acc < 0) {
                    out.ops.push_back({OP_GEMM_SET, where[vid], where[vid], c});
                    acc = vid;
                }
                else {
                    make_resident(acc, vid);
                    out.ops.push_back({OP_GEMM_MUL, where[acc], where[vid], c});
                    release(vid);
                }
            }
            else if (acc < 0) {
                acc = new_vector();
                int s = acquire(-1, -1);
                where[acc] = s;
                owner[s] = acc;
                out.ops.push_back({OP_LEAF_SET, s, 0, c});
            }
            else {
                make_resident(acc, -1);
                out.ops.push_back({OP_LEAF_MUL, where[acc], 0, c});
            }
        }
        make_resident(acc, -1);
        out.ops.push_back({OP_RESCALE, where[acc], 0, v});
        return acc;
    }
};

// Copies and validates the caller's tree; returns an error text or nullptr.
const char* import_tree(HostTree& t, const cafe_b200_tree* tree, int n_leaves)
{
    const int nn = tree->n_nodes;
    t.n_nodes = nn;
    t.parent.assign(tree->parent, tree->parent + nn);
    t.child_offset.assign(tree->child_offset, tree->child_offset + nn + 1);
    t.child_list.assign(tree->child_list, tree->child_list + (nn - 1));
    t.leaf_col.assign(tree->leaf_col, tree->leaf_col + nn);
    t.lambda_index.assign(tree->lambda_index, tree->lambda_index + nn);
    t.branch.assign(tree->branch, tree->branch + nn);
    t.branch_key.resize(nn);
    t.n_internal = 0;
    t.n_lambdas = 1;
    if (t.child_offset[0] != 0 || t.child_offset[nn] != nn - 1) return "child_offset does not describe n_nodes-1 edges";
    int leaves = 0;
    for (int v = 0; v < nn; ++v) {
        if ((t.parent[v] < 0) != (v == nn - 1)) return "the root must be the last node and the only one without parent";
        if (v < nn - 1 && (t.parent[v] <= v || t.parent[v] >= nn)) return "children must precede their parent";
        if (t.child_offset[v + 1] < t.child_offset[v]) return "child_offset not monotone";
        for (int e = t.child_offset[v]; e < t.child_offset[v + 1]; ++e)
            if (t.child_list[e] < 0 || t.child_list[e] >= v || t.parent[t.child_list[e]] != v) return "child_list inconsistent with parent";
        if (t.is_leaf(v)) {
            ++leaves;
            if (n_leaves >= 0 && (t.leaf_col[v] < 0 || t.leaf_col[v] >= n_leaves)) return "leaf_col out of range";
        }
        else t.n_internal++;
        if (t.lambda_index[v] < 0) return "negative lambda index";
        t.n_lambdas = std::max(t.n_lambdas, t.lambda_index[v] + 1);
        t.branch_key[v] = (long)(t.branch[v] * 1000);
    }
    if (n_leaves >= 0 && leaves != n_leaves) return "n_leaves does not match the tree";
    if (t.is_leaf(nn - 1)) return "the root is a leaf";
    return nullptr;
}

// Peephole over the schedule (pruning without error model): a leaf sibling that directly follows is
// folded into the producing op.  The products formed are the same two-operand products in the same
// order, so results are bit-identical to the unfused schedule.
//   LEAF_SET(a,l1) LEAF_MUL(a,l2)  -> LEAF_SET2(a,l1,l2)
//   GEMM_SET(a,c)  LEAF_MUL(a,l)   -> GEMM_SET_LEAF(a,c,l)
// GEMM_MUL is never fused: (f1*f2)*leaf must not become f1*(f2*leaf).
std::vector<FusedOp> fuse_schedule(const std::vector<Op>& ops, bool fuse)
{
    std::vector<FusedOp> out;
    for (size_t i = 0; i < ops.size(); ++i) {
        const Op& o = ops[i];
        if (fuse && i + 1 < ops.size() && ops[i + 1].type == OP_LEAF_MUL && ops[i + 1].a == o.a) {
            if (o.type == OP_LEAF_SET) { out.push_back({OP_LEAF_SET2, o.a, 0, o.node, ops[i + 1].node}); ++i; continue; }
            if (o.type == OP_GEMM_SET) { out.push_back({OP_GEMM_SET_LEAF, o.a, o.b, o.node, ops[i + 1].node}); ++i; continue; }
        }
        out.push_back({o.type, o.a, o.b, o.node, -1});
    }
    return out;
}

template <typename T>
cudaError_t dev_alloc(T** p, size_t n, bool zero = false)
{
    cudaError_t e = cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess && zero) e = cudaMemset(*p, 0, std::max<size_t>(n, 1) * sizeof(T));
    return e;
}

}  // namespace

struct cafe_b200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    int sm_count = 0;
    int smem_optin = 0;
    HostTree tree;
    int n_leaves = 0, mf = 0, mrf = 0, n = 0, mb = 0, nr = 0, kpanels = 0, n_kchunks = 0;
    int64_t n_families = 0;
    int64_t n_tiles = 0;
    int max_count = 0;
    Schedule sched;                     // reconstruction kernel: op list for n_slots slots
    Schedule psched;                    // pruning kernel: op list for prune_slots slots (unfused form: used with an error model)
    std::vector<FusedOp> fused;         // psched with leaf siblings fused: pruning without error model
    int n_slots = 0, hw_slots = 0;              // reconstruction kernel (32-family tiles)
    int prune_slots = 0, prune_hw_slots = 0;    // pruning kernel (n_groups x 16-family tiles)
    int n_groups = 2;                   // consumer groups of the pruning kernel (3 when shared memory allows)
    int n_stages = 4;                   // pruning ring depth
    int rescale = 0;
    int cap_k = 0;                      // categories the k-dependent buffers are sized for
    size_t mp_stride = 0, mt_stride = 0;
    // device buffers
    int32_t* d_counts = nullptr;
    Op* d_ops = nullptr;
    POp* d_pops = nullptr;              // [cap_k][pops_cap] per-category resolved pruning ops
    int pops_cap = 0;
    int n_pops = 0;                     // ops per category in the last staged evaluation
    int* d_leaf_col = nullptr;
    int* d_parent = nullptr;
    int* d_child_offset = nullptr;
    int* d_child_list = nullptr;
    int* d_mat_of = nullptr;
    double* d_mp = nullptr;
    double* d_mt = nullptr;
    KeyParams* d_keys = nullptr;
    double* d_powc = nullptr;
    double* d_lgamma = nullptr;
    double* d_err = nullptr;
    int err_rows = 0, err_ndev = 0;
    double* d_prior = nullptr;
    double* d_logprior = nullptr;
    double* d_catprobs = nullptr;
    double* d_cat_lk = nullptr;
    uint8_t* d_fail = nullptr;
    double* d_family_lnl = nullptr;
    uint8_t* d_family_fail = nullptr;
    double* d_partial = nullptr;
    double* d_result = nullptr;
    double* d_scratch = nullptr;        // reconstruction spill area
    double* d_pscratch = nullptr;       // pruning spill area [SMs][n_spill][tile families x LDV]
    int* d_pscratch_exp = nullptr;
    // pinned staging
    unsigned char* h_stage = nullptr;
    size_t h_stage_bytes = 0;
    double* h_result = nullptr;
    cudaEvent_t staged = nullptr;       // H2D copies of the last call have consumed h_stage
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    bool ev_valid[5] = {false, false, false, false, false};
    int64_t launches = 0;
    int64_t evals = 0;
    std::string error;
};

namespace {

int fail(cafe_b200_ctx* ctx, int code, const std::string& msg)
{
    if (ctx) ctx->error = msg;
    else g_create_error = msg;
    return code;
}

#define CUDA_TRY(ctx, call)                                                                              \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess) return fail(ctx, CAFE_B200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

int ensure_category_buffers(cafe_b200_ctx* c, int k)
{
    if (k <= c->cap_k) return CAFE_B200_OK;
    cudaFree(c->d_mat_of); cudaFree(c->d_mp); cudaFree(c->d_mt); cudaFree(c->d_keys); cudaFree(c->d_powc);
    cudaFree(c->d_cat_lk); cudaFree(c->d_fail); cudaFree(c->d_catprobs); cudaFree(c->d_pops);
    c->d_pops = nullptr;
    c->d_mat_of = nullptr; c->d_mp = c->d_mt = nullptr; c->d_keys = nullptr; c->d_powc = nullptr;
    c->d_cat_lk = nullptr; c->d_fail = nullptr; c->d_catprobs = nullptr;
    c->cap_k = 0;
    const size_t keys = (size_t)k * c->tree.n_nodes;
    CUDA_TRY(c, dev_alloc(&c->d_mat_of, keys));
    CUDA_TRY(c, dev_alloc(&c->d_mp, keys * c->mp_stride, true));
    CUDA_TRY(c, dev_alloc(&c->d_mt, keys * c->mt_stride, true));
    CUDA_TRY(c, dev_alloc(&c->d_keys, keys));
    CUDA_TRY(c, dev_alloc(&c->d_powc, keys * c->n));
    CUDA_TRY(c, dev_alloc(&c->d_cat_lk, (size_t)c->n_families * k));
    CUDA_TRY(c, dev_alloc(&c->d_fail, (size_t)c->n_families * k, true));
    CUDA_TRY(c, dev_alloc(&c->d_catprobs, (size_t)k));
    c->pops_cap = (int)c->psched.ops.size();
    CUDA_TRY(c, dev_alloc(&c->d_pops, (size_t)k * c->pops_cap));
    // staging: keys + powc + mat_of + prior + logprior + catprobs + per-category ops
    size_t need = keys * sizeof(KeyParams) + keys * c->n * sizeof(double) + keys * sizeof(int) + (2 * (size_t)c->n + k + 64) * sizeof(double)
                  + (size_t)k * c->pops_cap * sizeof(POp) + 64;
    if (need > c->h_stage_bytes) {
        if (c->h_stage) cudaFreeHost(c->h_stage);
        c->h_stage = nullptr;
        CUDA_TRY(c, cudaMallocHost((void**)&c->h_stage, need));
        c->h_stage_bytes = need;
    }
    c->cap_k = k;
    return CAFE_B200_OK;
}

// Quantise keys, de-duplicate, stage per-key scalars and launch the matrix builder.
int stage_and_build(cafe_b200_ctx* c, const double* lambdas, int n_lambdas, int k, const double* cat_probs, const double* prior,
                    int n_prior)
{
    if (n_lambdas < c->tree.n_lambdas) return fail(c, CAFE_B200_ERR_ARG, "n_lambdas smaller than the tree's lambda indices");
    int rc = ensure_category_buffers(c, k);
    if (rc) return rc;
    CUDA_TRY(c, cudaEventSynchronize(c->staged));
    const HostTree& t = c->tree;
    const size_t slots = (size_t)k * t.n_nodes;
    unsigned char* h = c->h_stage;
    KeyParams* h_keys = reinterpret_cast<KeyParams*>(h); h += slots * sizeof(KeyParams);
    double* h_powc = reinterpret_cast<double*>(h); h += slots * c->n * sizeof(double);
    double* h_prior = reinterpret_cast<double*>(h); h += (size_t)c->n * sizeof(double);
    double* h_logprior = reinterpret_cast<double*>(h); h += (size_t)c->n * sizeof(double);
    double* h_cat = reinterpret_cast<double*>(h); h += ((size_t)k + 8) * sizeof(double);
    int* h_mat_of = reinterpret_cast<int*>(h);

    std::map<std::pair<long, long>, int> seen;
    int n_keys = 0;
    for (int cat = 0; cat < k; ++cat) {
        for (int v = 0; v < t.n_nodes; ++v) {
            if (t.parent[v] < 0) { h_mat_of[cat * t.n_nodes + v] = 0; continue; }
            const double lam = lambdas[(size_t)cat * n_lambdas + t.lambda_index[v]];
            const long kl = (long)(lam * 1000000000);                        // src/matrix_cache.h:49
            const long kt = t.branch_key[v];
            auto it = seen.find({kl, kt});
            if (it == seen.end()) {
                const double lq = double(kl) / 1000000000.0;                 // src/matrix_cache.h:55-57
                const double tq = double(kt) / 1000.0;                       // src/matrix_cache.h:58-60
                const double alpha = lq * tq / (1 + lq * tq);                // src/probability.cpp:150
                const double coeff = 1 - 2 * alpha;                          // src/probability.cpp:151
                KeyParams kp;
                kp.saturated = (1 - 2 * alpha) < 0 ? 1 : 0;                  // src/matrix_cache.cpp:115-119
                kp.computable = (coeff > 0 && coeff != 1) ? 1 : 0;           // src/probability.cpp:154
                kp.log_alpha = std::log(alpha);
                kp.coeff = coeff;
                h_keys[n_keys] = kp;
                double* pw = h_powc + (size_t)n_keys * c->n;
                for (int j = 0; j < c->n; ++j) pw[j] = std::pow(coeff, (double)j);   // src/probability.cpp:125
                it = seen.emplace(std::make_pair(kl, kt), n_keys++).first;
            }
            h_mat_of[cat * t.n_nodes + v] = it->second;
        }
    }
    for (int j = 0; j < c->n; ++j) {
        const double pj = (prior && j < n_prior) ? prior[j] : 0.0;
        h_prior[j] = pj;
        h_logprior[j] = std::log(pj);                                        // src/base_model.cpp:98
    }
    for (int cat = 0; cat < k; ++cat) h_cat[cat] = cat_probs ? cat_probs[cat] : 1.0;
    // per-category pruning ops with matrix slots and count columns resolved
    POp* h_pops = reinterpret_cast<POp*>((reinterpret_cast<uintptr_t>(h_mat_of + slots) + 15) & ~uintptr_t(15));   // POp is 16-byte aligned
    if ((int)c->psched.ops.size() > c->pops_cap) return fail(c, CAFE_B200_ERR_ARG, "schedule grew after the category buffers were sized");
    const std::vector<FusedOp> plain = c->d_err ? fuse_schedule(c->psched.ops, false) : std::vector<FusedOp>();
    const std::vector<FusedOp>& fo = c->d_err ? plain : c->fused;
    c->n_pops = (int)fo.size();
    for (int cat = 0; cat < k; ++cat)
        for (int o = 0; o < c->n_pops; ++o) {
            POp q;
            q.type = fo[o].type; q.a = fo[o].a; q.b = fo[o].b; q.node = fo[o].node;
            q.mat = h_mat_of[cat * t.n_nodes + fo[o].node];
            q.col = t.leaf_col[fo[o].node];
            q.mat2 = fo[o].node2 >= 0 ? h_mat_of[cat * t.n_nodes + fo[o].node2] : 0;
            q.col2 = fo[o].node2 >= 0 ? t.leaf_col[fo[o].node2] : 0;
            h_pops[(size_t)cat * c->n_pops + o] = q;
        }

    cudaStream_t s = c->stream;
    CUDA_TRY(c, cudaMemcpyAsync(c->d_keys, h_keys, (size_t)n_keys * sizeof(KeyParams), cudaMemcpyHostToDevice, s));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_powc, h_powc, (size_t)n_keys * c->n * sizeof(double), cudaMemcpyHostToDevice, s));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_mat_of, h_mat_of, slots * sizeof(int), cudaMemcpyHostToDevice, s));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_prior, h_prior, (size_t)c->n * sizeof(double), cudaMemcpyHostToDevice, s));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_logprior, h_logprior, (size_t)c->n * sizeof(double), cudaMemcpyHostToDevice, s));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_catprobs, h_cat, (size_t)k * sizeof(double), cudaMemcpyHostToDevice, s));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_pops, h_pops, (size_t)k * c->n_pops * sizeof(POp), cudaMemcpyHostToDevice, s));
    CUDA_TRY(c, cudaEventRecord(c->staged, s));

    CUDA_TRY(c, cudaEventRecord(c->ev[0], s));
    MatrixBuildParams mp;
    mp.n = c->n; mp.mf = c->mf; mp.nr = c->nr; mp.n_keys = n_keys;
    mp.keys = c->d_keys; mp.powc = c->d_powc; mp.lgamma_tab = c->d_lgamma;
    mp.mp = c->d_mp; mp.mt = c->d_mt; mp.mp_stride = c->mp_stride; mp.mt_stride = c->mt_stride;
    const int entries = ((c->n + 31) / 32) * 32 * (c->mf + 1);
    int bx = (entries + MB_THREADS - 1) / MB_THREADS;
    // keep the whole launch near a few waves: many keys -> fewer blocks per key (grid-stride inside)
    const int target = std::max(1, (8 * c->sm_count + n_keys - 1) / n_keys);
    bx = std::max(1, std::min(bx, target));
    dim3 grid(bx, n_keys);
    const size_t smem = (LGAMMA_TABLE + c->n) * sizeof(double);
    bd_matrix_kernel<<<grid, MB_THREADS, smem, s>>>(mp);
    CUDA_TRY(c, cudaGetLastError());
    c->launches++;
    CUDA_TRY(c, cudaEventRecord(c->ev[1], s));
    c->ev_valid[0] = c->ev_valid[1] = true;
    return CAFE_B200_OK;
}

template <int MB, int NG>
int launch_prune_mbg(cafe_b200_ctx* c, const PruneParams& p)
{
    using L = PruneSmem<MB, NG>;
    const int smem = L::total_bytes(c->prune_slots, c->n_stages);
    CUDA_TRY(c, cudaFuncSetAttribute(prune_kernel<MB, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int64_t items = p.n_tiles * p.n_categories;        // one item = one tile of NG x 16 families of one category
    const int grid = (int)std::min<int64_t>(items, c->sm_count);
    prune_kernel<MB, NG><<<grid, prune_threads(NG), smem, c->stream>>>(p);
    CUDA_TRY(c, cudaGetLastError());
    c->launches++;
    return CAFE_B200_OK;
}

template <int MB>
int launch_prune_mb(cafe_b200_ctx* c, const PruneParams& p)
{
    return c->n_groups == 3 ? launch_prune_mbg<MB, 3>(c, p) : launch_prune_mbg<MB, 2>(c, p);
}

int launch_prune(cafe_b200_ctx* c, int k, int mode, double* root_out)
{
    if (c->n_families == 0) return CAFE_B200_OK;
    const int pft = c->n_groups * GFT;
    PruneParams p;
    memset(&p, 0, sizeof(p));
    p.n_families = c->n_families; p.n_leaves = c->n_leaves; p.n_nodes = c->tree.n_nodes; p.n_categories = k;
    p.mf = c->mf; p.mrf = c->mrf; p.n_ops = c->n_pops; p.n_kchunks = c->n_kchunks; p.mode = mode;
    p.rescale = c->rescale; p.n_spill = std::max(1, c->psched.n_spill); p.err_rows = c->err_rows; p.err_ndev = c->err_ndev;
    p.counts_in_smem = (pft * c->n_leaves * 2 <= PRUNE_CNT_CAP_BYTES) ? 1 : 0;
    p.n_slots = c->prune_slots; p.n_tiles = (c->n_families + pft - 1) / pft;
    // c->n_stages counts 10 KB chunks of ring memory; the kernel's stages hold CPS chunks each
    p.n_stages = c->n_stages / CPS; p.stage_shift = p.n_stages == 4 ? 2 : (p.n_stages == 2 ? 1 : 0);
    p.ops = c->d_pops; p.counts = c->d_counts;
    p.mp = c->d_mp; p.mt = c->d_mt; p.mp_stride = c->mp_stride; p.mt_stride = c->mt_stride;
    p.err = c->d_err; p.prior = c->d_prior; p.logprior = c->d_logprior; p.cat_probs = c->d_catprobs;
    p.scratch = c->d_pscratch; p.scratch_exp = c->d_pscratch_exp;
    p.cat_lk = c->d_cat_lk; p.fail = c->d_fail; p.root_out = root_out;
    switch (c->mb) {
    case 1: return launch_prune_mb<1>(c, p);
    case 2: return launch_prune_mb<2>(c, p);
    case 3: return launch_prune_mb<3>(c, p);
    case 4: return launch_prune_mb<4>(c, p);
    case 5: return launch_prune_mb<5>(c, p);
    case 6: return launch_prune_mb<6>(c, p);
    case 7: return launch_prune_mb<7>(c, p);
    case 8: return launch_prune_mb<8>(c, p);
    }
    return fail(c, CAFE_B200_ERR_LIMIT, "matrix size not supported");
}

template <int MB>
int launch_pupko_mb(cafe_b200_ctx* c, const PupkoParams& p)
{
    using L = PupkoSmem<MB>;
    const int smem = L::total_bytes(c->n_slots);
    CUDA_TRY(c, cudaFuncSetAttribute(pupko_kernel<MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int64_t items = p.n_tiles * p.n_categories;
    const int grid = (int)std::min<int64_t>(items, c->sm_count);
    pupko_kernel<MB><<<grid, PRUNE_THREADS, smem, c->stream>>>(p);
    CUDA_TRY(c, cudaGetLastError());
    c->launches++;
    return CAFE_B200_OK;
}

int launch_pupko(cafe_b200_ctx* c, int k, int32_t* states_host)
{
    if (c->n_families == 0) return CAFE_B200_OK;
    const HostTree& t = c->tree;
    std::vector<int> internal_idx(t.n_nodes, -1);
    int ni = 0;
    for (int v = 0; v < t.n_nodes; ++v) if (!t.is_leaf(v)) internal_idx[v] = ni++;
    int* d_internal = nullptr;
    uint8_t* d_ctab = nullptr;
    int32_t* d_states = nullptr;
    const size_t n_states = (size_t)c->n_families * k * t.n_internal;
    const int grid = (int)std::min<int64_t>(c->n_tiles * k, c->sm_count);
    int rc = CAFE_B200_OK;
    cudaError_t e = dev_alloc(&d_internal, (size_t)t.n_nodes);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_internal, internal_idx.data(), t.n_nodes * sizeof(int), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = dev_alloc(&d_ctab, (size_t)grid * t.n_nodes * FT * c->nr);
    if (e == cudaSuccess) e = dev_alloc(&d_states, n_states);
    if (e == cudaSuccess) {
        PupkoParams p;
        memset(&p, 0, sizeof(p));
        p.n_families = c->n_families; p.n_leaves = c->n_leaves; p.n_nodes = t.n_nodes; p.n_internal = t.n_internal; p.n_categories = k;
        p.mf = c->mf; p.mrf = c->mrf; p.n_ops = (int)c->sched.ops.size(); p.n_kchunks = c->n_kchunks;
        p.n_spill = std::max(1, c->sched.n_spill); p.n_slots = c->n_slots;
        p.counts_in_smem = (FT * c->n_leaves * 2 <= CNT_CAP_BYTES) ? 1 : 0;
        p.n_tiles = c->n_tiles; p.ops = c->d_ops; p.counts = c->d_counts; p.leaf_col = c->d_leaf_col; p.parent = c->d_parent;
        p.internal_idx = d_internal; p.mat_of = c->d_mat_of; p.mt = c->d_mt; p.mt_stride = c->mt_stride; p.prior = c->d_prior;
        p.scratch = c->d_scratch; p.ctab = d_ctab; p.states = d_states;
        switch (c->mb) {
        case 1: rc = launch_pupko_mb<1>(c, p); break;
        case 2: rc = launch_pupko_mb<2>(c, p); break;
        case 3: rc = launch_pupko_mb<3>(c, p); break;
        case 4: rc = launch_pupko_mb<4>(c, p); break;
        case 5: rc = launch_pupko_mb<5>(c, p); break;
        case 6: rc = launch_pupko_mb<6>(c, p); break;
        case 7: rc = launch_pupko_mb<7>(c, p); break;
        default: rc = launch_pupko_mb<8>(c, p); break;
        }
        if (rc == CAFE_B200_OK) {
            e = cudaEventRecord(c->ev[4], c->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(states_host, d_states, n_states * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        }
    }
    cudaStreamSynchronize(c->stream);
    cudaFree(d_internal); cudaFree(d_ctab); cudaFree(d_states);
    if (e != cudaSuccess) return fail(c, CAFE_B200_ERR_CUDA, std::string("pupko: ") + cudaGetErrorString(e));
    return rc;
}

// Shared-memory budget of the two tree-walking kernels.  Pruning: two consumer groups (32-family tiles) with three
// vector slots and the deepest ring that fits (8 stages, else 4), two slots as a last resort.  A three-group layout
// (48-family tiles, two slots; CAFE_B200_GROUPS=3) exists for experiments: the third MMA warp per sub-partition helps
// an isolated K loop (scripts/kloop_mix.cu: 0.81 vs 0.76 of the DMMA rate) but in the full kernel the 128-register
// cap of 448 threads and the extra spills of a two-slot schedule cost more (measured 0.685 vs 0.706).
template <int MB>
bool plan_shared_memory_mb(cafe_b200_ctx* c)
{
    const int lim = c->smem_optin;
    c->hw_slots = std::min(MAX_SLOTS, (lim - PupkoSmem<MB>::total_bytes(0)) / PupkoSmem<MB>::SLOT_BYTES);
    c->prune_hw_slots = 0;
    const char* e = getenv("CAFE_B200_GROUPS");
    if (e && atoi(e) == 3)
        for (int stages : {8, 4}) {
            const int sl = PruneSmem<MB, 3>::max_slots(lim, stages);
            if (sl >= 2) { c->n_groups = 3; c->n_stages = stages; c->prune_hw_slots = sl; break; }
        }
    if (!c->prune_hw_slots)
        for (int stages : {8, 4}) {
            const int sl = PruneSmem<MB, 2>::max_slots(lim, stages);
            if (sl >= 3 || (stages == 4 && sl >= 2)) { c->n_groups = 2; c->n_stages = stages; c->prune_hw_slots = sl; break; }
        }
    c->n_slots = c->hw_slots;
    c->prune_slots = c->prune_hw_slots;
    return c->hw_slots >= 2 && c->prune_hw_slots >= 2;
}

bool plan_shared_memory(cafe_b200_ctx* c)
{
    switch (c->mb) {
    case 1: return plan_shared_memory_mb<1>(c);
    case 2: return plan_shared_memory_mb<2>(c);
    case 3: return plan_shared_memory_mb<3>(c);
    case 4: return plan_shared_memory_mb<4>(c);
    case 5: return plan_shared_memory_mb<5>(c);
    case 6: return plan_shared_memory_mb<6>(c);
    case 7: return plan_shared_memory_mb<7>(c);
    default: return plan_shared_memory_mb<8>(c);
    }
}

// (Re)build the op lists (reconstruction: n_slots, pruning: prune_slots) and size the spill scratch for them.
int upload_schedule(cafe_b200_ctx* c)
{
    c->sched = ScheduleBuilder(c->tree, c->n_slots).build();
    c->psched = ScheduleBuilder(c->tree, c->prune_slots).build();
    c->fused = fuse_schedule(c->psched.ops, true);
    if ((int)c->psched.ops.size() > c->pops_cap) c->cap_k = 0;       // force the per-category op buffers to be re-sized
    cudaFree(c->d_ops); cudaFree(c->d_scratch); cudaFree(c->d_pscratch); cudaFree(c->d_pscratch_exp);
    c->d_ops = nullptr; c->d_scratch = nullptr; c->d_pscratch = nullptr; c->d_pscratch_exp = nullptr;
    CUDA_TRY(c, dev_alloc(&c->d_ops, c->sched.ops.size()));
    CUDA_TRY(c, cudaMemcpy(c->d_ops, c->sched.ops.data(), c->sched.ops.size() * sizeof(Op), cudaMemcpyHostToDevice));
    const size_t ldv = ldv_of(c->mb);
    CUDA_TRY(c, dev_alloc(&c->d_scratch, (size_t)c->sm_count * std::max(1, c->sched.n_spill) * FT * ldv, true));
    const size_t pft = (size_t)c->n_groups * GFT, pnsp = (size_t)std::max(1, c->psched.n_spill);
    CUDA_TRY(c, dev_alloc(&c->d_pscratch, (size_t)c->sm_count * pnsp * pft * ldv, true));
    CUDA_TRY(c, dev_alloc(&c->d_pscratch_exp, (size_t)c->sm_count * pnsp * pft, true));
    return CAFE_B200_OK;
}

int check_counts(cafe_b200_ctx* c)
{
    int hi = c->max_count;
    if (c->d_err) {
        if (c->max_count >= c->err_rows) return fail(c, CAFE_B200_ERR_COUNT_RANGE, "a leaf count has no error-model row");
        hi += (c->err_ndev - 1) / 2;
    }
    if (hi > c->mf) return fail(c, CAFE_B200_ERR_COUNT_RANGE, "a leaf count (plus error-model deviation) exceeds max_family_size");
    return CAFE_B200_OK;
}

int run_eval(cafe_b200_ctx* c, const double* lambdas, int n_lambdas, const double* cat_probs, int k, const double* prior, int mode,
             double* result_device)
{
    if (!c || !lambdas || k < 1 || n_lambdas < 1 || (mode != CAFE_B200_BASE_LOGMAX && mode != CAFE_B200_GAMMA_LINSUM))
        return fail(c, CAFE_B200_ERR_ARG, "bad argument to eval");
    if (mode == CAFE_B200_BASE_LOGMAX && k != 1) return fail(c, CAFE_B200_ERR_ARG, "base mode takes exactly one category");
    if (!prior) return fail(c, CAFE_B200_ERR_ARG, "prior is required");
    CUDA_TRY(c, cudaSetDevice(c->device));
    int rc = check_counts(c);
    if (rc) return rc;
    rc = stage_and_build(c, lambdas, n_lambdas, k, cat_probs, prior, c->mrf);
    if (rc) return rc;
    rc = launch_prune(c, k, mode, nullptr);
    if (rc) return rc;
    cudaStream_t s = c->stream;
    CUDA_TRY(c, cudaEventRecord(c->ev[2], s));
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((c->n_families + RED_THREADS - 1) / RED_THREADS, 1024));
    finalize_kernel<<<blocks, RED_THREADS, 0, s>>>(c->n_families, k, mode, c->d_cat_lk, c->d_fail, c->d_family_lnl, c->d_family_fail, c->d_partial);
    CUDA_TRY(c, cudaGetLastError());
    final_sum_kernel<<<1, RED_THREADS, 0, s>>>(blocks, c->d_partial, result_device);
    CUDA_TRY(c, cudaGetLastError());
    c->launches += 2;
    CUDA_TRY(c, cudaEventRecord(c->ev[3], s));
    c->ev_valid[2] = c->ev_valid[3] = true;
    c->ev_valid[4] = false;
    c->evals++;
    return CAFE_B200_OK;
}

}  // namespace

extern "C" {

int cafe_b200_abi_version(void) { return CAFE_B200_ABI_VERSION; }

void cafe_b200_get_limits(cafe_b200_limits* out)
{
    if (!out) return;
    out->max_matrix_size = 32 * MAX_MB;
    out->max_categories = 64;
    out->max_nodes = 1 << 20;
    out->families_per_tile = FT;
}

int cafe_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* cafe_b200_last_error(const cafe_b200_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

void cafe_b200_destroy(cafe_b200_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->d_counts); cudaFree(c->d_ops); cudaFree(c->d_leaf_col); cudaFree(c->d_parent); cudaFree(c->d_child_offset);
    cudaFree(c->d_child_list); cudaFree(c->d_mat_of); cudaFree(c->d_mp); cudaFree(c->d_mt); cudaFree(c->d_keys); cudaFree(c->d_powc);
    cudaFree(c->d_lgamma); cudaFree(c->d_err); cudaFree(c->d_prior); cudaFree(c->d_logprior); cudaFree(c->d_catprobs);
    cudaFree(c->d_pops);
    cudaFree(c->d_cat_lk); cudaFree(c->d_fail); cudaFree(c->d_family_lnl); cudaFree(c->d_family_fail); cudaFree(c->d_partial);
    cudaFree(c->d_result); cudaFree(c->d_scratch); cudaFree(c->d_pscratch); cudaFree(c->d_pscratch_exp);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->h_result) cudaFreeHost(c->h_result);
    if (c->staged) cudaEventDestroy(c->staged);
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

int cafe_b200_create(cafe_b200_ctx** out, const cafe_b200_tree* tree, const int32_t* leaf_counts, int64_t n_families, int n_leaves,
                     int max_family_size, int max_root_family_size, int device)
{
    if (!out || !tree || n_families < 0 || n_leaves < 1 || max_family_size < 1 || max_root_family_size < 1 || (n_families > 0 && !leaf_counts))
        return fail(nullptr, CAFE_B200_ERR_ARG, "bad argument to create");
    *out = nullptr;
    const int nn = tree->n_nodes;
    if (nn < 2 || !tree->parent || !tree->child_offset || !tree->child_list || !tree->leaf_col || !tree->branch || !tree->lambda_index)
        return fail(nullptr, CAFE_B200_ERR_ARG, "incomplete tree");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, CAFE_B200_ERR_CUDA, "no CUDA device: this engine has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(nullptr, CAFE_B200_ERR_ARG, "device ordinal out of range");

    cafe_b200_ctx* c = new cafe_b200_ctx();
    c->device = device;
    HostTree& t = c->tree;
    {
        const char* why = import_tree(t, tree, n_leaves);
        if (why) { g_create_error = why; delete c; return CAFE_B200_ERR_ARG; }
    }
    c->n_leaves = n_leaves; c->mf = max_family_size; c->mrf = max_root_family_size;
    c->n = std::max(c->mf, c->mrf) + 1;                                      // src/base_model.cpp:77
    c->n_families = n_families;
    c->n_tiles = (n_families + FT - 1) / FT;
    c->mb = (c->n + 31) / 32;
    if (c->mb > MAX_MB || 2 * c->n + 2 > LGAMMA_TABLE) {
        g_create_error = "matrix size beyond this build (max 256)";
        delete c;
        return CAFE_B200_ERR_LIMIT;
    }
    c->nr = nr_of(c->mb);
    c->kpanels = ((c->mf + 1 + 3) / 4 + PPS - 1) / PPS * PPS;
    c->n_kchunks = c->kpanels / PPS;
    c->mp_stride = (size_t)c->kpanels * c->nr * 4;
    c->mt_stride = (size_t)c->kpanels * 4 * c->nr;      // columns padded to whole ring stages (zeros)
    for (int64_t i = 0; i < n_families * n_leaves; ++i) {
        if (leaf_counts[i] < 0) { g_create_error = "negative leaf count"; delete c; return CAFE_B200_ERR_COUNT_RANGE; }
        c->max_count = std::max(c->max_count, (int)leaf_counts[i]);
    }
    if (c->max_count > c->mf) { g_create_error = "a leaf count exceeds max_family_size"; delete c; return CAFE_B200_ERR_COUNT_RANGE; }

#define CREATE_TRY(call)                                                                      \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            g_create_error = std::string(#call) + ": " + cudaGetErrorString(e__);             \
            cafe_b200_destroy(c);                                                             \
            return CAFE_B200_ERR_CUDA;                                                        \
        }                                                                                     \
    } while (0)

    CREATE_TRY(cudaSetDevice(device));
    CREATE_TRY(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    CREATE_TRY(cudaDeviceGetAttribute(&c->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    CREATE_TRY(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    CREATE_TRY(cudaEventCreateWithFlags(&c->staged, cudaEventDisableTiming));
    CREATE_TRY(cudaEventRecord(c->staged, c->stream));
    for (auto& e : c->ev) CREATE_TRY(cudaEventCreate(&e));

    if (!plan_shared_memory(c)) { g_create_error = "not enough shared memory for two vector slots"; cafe_b200_destroy(c); return CAFE_B200_ERR_LIMIT; }
    if (upload_schedule(c) != CAFE_B200_OK) { g_create_error = c->error; cafe_b200_destroy(c); return CAFE_B200_ERR_CUDA; }

    CREATE_TRY(dev_alloc(&c->d_counts, (size_t)n_families * n_leaves));
    if (n_families) CREATE_TRY(cudaMemcpy(c->d_counts, leaf_counts, (size_t)n_families * n_leaves * sizeof(int32_t), cudaMemcpyHostToDevice));
    CREATE_TRY(dev_alloc(&c->d_leaf_col, (size_t)nn));
    CREATE_TRY(cudaMemcpy(c->d_leaf_col, t.leaf_col.data(), nn * sizeof(int), cudaMemcpyHostToDevice));
    CREATE_TRY(dev_alloc(&c->d_parent, (size_t)nn));
    CREATE_TRY(cudaMemcpy(c->d_parent, t.parent.data(), nn * sizeof(int), cudaMemcpyHostToDevice));
    CREATE_TRY(dev_alloc(&c->d_child_offset, (size_t)nn + 1));
    CREATE_TRY(cudaMemcpy(c->d_child_offset, t.child_offset.data(), (nn + 1) * sizeof(int), cudaMemcpyHostToDevice));
    CREATE_TRY(dev_alloc(&c->d_child_list, (size_t)nn));
    CREATE_TRY(cudaMemcpy(c->d_child_list, t.child_list.data(), (nn - 1) * sizeof(int), cudaMemcpyHostToDevice));
    std::vector<double> lg(LGAMMA_TABLE);
    for (int i = 0; i < LGAMMA_TABLE; ++i) lg[i] = lgamma((double)i);        // src/probability.cpp:66-72
    CREATE_TRY(dev_alloc(&c->d_lgamma, (size_t)LGAMMA_TABLE));
    CREATE_TRY(cudaMemcpy(c->d_lgamma, lg.data(), LGAMMA_TABLE * sizeof(double), cudaMemcpyHostToDevice));
    CREATE_TRY(dev_alloc(&c->d_prior, (size_t)c->n));
    CREATE_TRY(dev_alloc(&c->d_logprior, (size_t)c->n));
    CREATE_TRY(dev_alloc(&c->d_family_lnl, (size_t)n_families));
    CREATE_TRY(dev_alloc(&c->d_family_fail, (size_t)n_families, true));
    CREATE_TRY(dev_alloc(&c->d_partial, (size_t)2 * 1024));
    CREATE_TRY(dev_alloc(&c->d_result, (size_t)2));
    CREATE_TRY(cudaMallocHost((void**)&c->h_result, 2 * sizeof(double)));
#undef CREATE_TRY
    *out = c;
    return CAFE_B200_OK;
}

int cafe_b200_set_families(cafe_b200_ctx* c, const int32_t* leaf_counts, int64_t n_families)
{
    if (!c || !leaf_counts || n_families != c->n_families) return fail(c, CAFE_B200_ERR_ARG, "set_families: shape must match create");
    const int64_t total = n_families * c->n_leaves;
    if (total == 0) return CAFE_B200_OK;
    CUDA_TRY(c, cudaSetDevice(c->device));
    // upload, then range-check on the device (a host pass over 10^8 counts costs more than the copy)
    int* d_range = reinterpret_cast<int*>(c->d_partial);      // scratch: [min, max]
    const int init[2] = {INT_MAX, INT_MIN};
    CUDA_TRY(c, cudaMemcpyAsync(d_range, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_counts, leaf_counts, (size_t)total * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((total / 4 + RED_THREADS - 1) / RED_THREADS, 8 * (int64_t)c->sm_count));
    count_range_kernel<<<blocks, RED_THREADS, 0, c->stream>>>(c->d_counts, total, d_range);
    CUDA_TRY(c, cudaGetLastError());
    c->launches++;
    int range[2] = {0, 0};
    CUDA_TRY(c, cudaMemcpyAsync(range, d_range, sizeof(range), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    // the context now holds the new matrix; out-of-range counts are refused here and again by every evaluation
    c->max_count = range[1];
    if (range[0] < 0) { c->max_count = c->mf + 1; return fail(c, CAFE_B200_ERR_COUNT_RANGE, "negative leaf count"); }
    if (range[1] > c->mf) return fail(c, CAFE_B200_ERR_COUNT_RANGE, "a leaf count exceeds max_family_size");
    return CAFE_B200_OK;
}

int cafe_b200_set_error_model(cafe_b200_ctx* c, const double* probs, int rows, int n_deviations)
{
    if (!c) return CAFE_B200_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->d_err);
    c->d_err = nullptr; c->err_rows = c->err_ndev = 0;
    if (!probs) return CAFE_B200_OK;
    if (rows < 1 || n_deviations < 1 || (n_deviations % 2) == 0) return fail(c, CAFE_B200_ERR_ARG, "error model: odd number of deviations required");
    CUDA_TRY(c, dev_alloc(&c->d_err, (size_t)rows * n_deviations));
    CUDA_TRY(c, cudaMemcpy(c->d_err, probs, (size_t)rows * n_deviations * sizeof(double), cudaMemcpyHostToDevice));
    c->err_rows = rows; c->err_ndev = n_deviations;
    return CAFE_B200_OK;
}

int cafe_b200_set_option(cafe_b200_ctx* c, int option, int value)
{
    if (!c) return CAFE_B200_ERR_ARG;
    if (option == CAFE_B200_OPT_RESCALE) { c->rescale = value ? 1 : 0; return CAFE_B200_OK; }
    if (option == CAFE_B200_OPT_MAX_SLOTS) {
        if (value < 2) return fail(c, CAFE_B200_ERR_ARG, "at least two slots are required");
        CUDA_TRY(c, cudaSetDevice(c->device));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        c->n_slots = std::min(value, c->hw_slots);
        c->prune_slots = std::min(value, c->prune_hw_slots);
        return upload_schedule(c);
    }
    return fail(c, CAFE_B200_ERR_ARG, "unknown option");
}

int cafe_b200_plan_schedule(const cafe_b200_tree* tree, int n_slots, int* ops_out, int cap, int* n_ops, int* n_spill)
{
    if (!tree || tree->n_nodes < 2 || n_slots < 2 || !n_ops) return CAFE_B200_ERR_ARG;
    HostTree t;
    const char* why = import_tree(t, tree, -1);
    if (why) { g_create_error = why; return CAFE_B200_ERR_ARG; }
    Schedule s = ScheduleBuilder(t, n_slots).build();
    *n_ops = (int)s.ops.size();
    if (n_spill) *n_spill = s.n_spill;
    if (ops_out) {
        if (cap < (int)s.ops.size()) return CAFE_B200_ERR_ARG;
        for (size_t i = 0; i < s.ops.size(); ++i) {
            ops_out[4 * i] = s.ops[i].type; ops_out[4 * i + 1] = s.ops[i].a; ops_out[4 * i + 2] = s.ops[i].b; ops_out[4 * i + 3] = s.ops[i].node;
        }
    }
    return CAFE_B200_OK;
}

int cafe_b200_set_stream(cafe_b200_ctx* c, void* cuda_stream)
{
    if (!c) return CAFE_B200_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->stream = (cudaStream_t)cuda_stream;          // NULL is the legacy default stream, as in the CUDA runtime
    CUDA_TRY(c, cudaEventRecord(c->staged, c->stream));
    return CAFE_B200_OK;
}

int cafe_b200_eval_device(cafe_b200_ctx* c, const double* lambdas, int n_lambdas, const double* cat_probs, int k, const double* prior,
                          int mode, double* result_device)
{
    if (!result_device) return fail(c, CAFE_B200_ERR_ARG, "result_device is null");
    return run_eval(c, lambdas, n_lambdas, cat_probs, k, prior, mode, result_device);
}

int cafe_b200_eval(cafe_b200_ctx* c, const double* lambdas, int n_lambdas, const double* cat_probs, int k, const double* prior, int mode,
                   double* neg_lnl, double* family_lnl, double* cat_lk, int64_t* n_failed, int64_t* failed_idx, int64_t failed_cap)
{
    if (!c || !neg_lnl) return fail(c, CAFE_B200_ERR_ARG, "neg_lnl is null");
    int rc = run_eval(c, lambdas, n_lambdas, cat_probs, k, prior, mode, c->d_result);
    if (rc) return rc;
    cudaStream_t s = c->stream;
    CUDA_TRY(c, cudaMemcpyAsync(c->h_result, c->d_result, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (family_lnl && c->n_families)
        CUDA_TRY(c, cudaMemcpyAsync(family_lnl, c->d_family_lnl, (size_t)c->n_families * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (cat_lk && c->n_families && mode == CAFE_B200_GAMMA_LINSUM)
        CUDA_TRY(c, cudaMemcpyAsync(cat_lk, c->d_cat_lk, (size_t)c->n_families * k * sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(c, cudaStreamSynchronize(s));
    if (c->n_families == 0) { c->h_result[0] = 0.0; c->h_result[1] = 0.0; }
    const int64_t nf = (int64_t)c->h_result[1];
    if (n_failed) *n_failed = nf;
    *neg_lnl = nf > 0 ? INFINITY : -c->h_result[0];                       // src/gamma_core.cpp:227-236 / src/base_model.cpp:107
    if (nf > 0 && failed_idx && failed_cap > 0) {
        std::vector<uint8_t> flags((size_t)c->n_families);
        CUDA_TRY(c, cudaMemcpy(flags.data(), c->d_family_fail, flags.size(), cudaMemcpyDeviceToHost));
        int64_t w = 0;
        for (int64_t i = 0; i < c->n_families && w < failed_cap; ++i)
            if (flags[i]) failed_idx[w++] = i;
    }
    return CAFE_B200_OK;
}

int cafe_b200_matrix_size(const cafe_b200_ctx* c) { return c ? c->n : 0; }

int cafe_b200_build_matrices(cafe_b200_ctx* c, const double* lambdas, int n_lambdas, int k, double* out)
{
    if (!c || !lambdas || !out || k < 1) return fail(c, CAFE_B200_ERR_ARG, "bad argument to build_matrices");
    CUDA_TRY(c, cudaSetDevice(c->device));
    std::vector<double> ones(c->n, 1.0);
    int rc = stage_and_build(c, lambdas, n_lambdas, k, nullptr, ones.data(), c->n);
    if (rc) return rc;
    const HostTree& t = c->tree;
    std::vector<int> mat_of((size_t)k * t.n_nodes);
    std::vector<double> mt(c->mt_stride);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    CUDA_TRY(c, cudaMemcpy(mat_of.data(), c->d_mat_of, mat_of.size() * sizeof(int), cudaMemcpyDeviceToHost));
    const int cols = c->mf + 1;
    for (int cat = 0; cat < k; ++cat)
        for (int v = 0; v < t.n_nodes; ++v) {
            double* dst = out + ((size_t)cat * t.n_nodes + v) * c->n * cols;
            if (t.parent[v] < 0) { std::fill(dst, dst + (size_t)c->n * cols, 0.0); continue; }
            CUDA_TRY(c, cudaMemcpy(mt.data(), c->d_mt + (size_t)mat_of[cat * t.n_nodes + v] * c->mt_stride, c->mt_stride * sizeof(double), cudaMemcpyDeviceToHost));
            for (int s = 0; s < c->n; ++s)
                for (int cc = 0; cc < cols; ++cc) dst[(size_t)s * cols + cc] = mt[(size_t)cc * c->nr + s];
        }
    return CAFE_B200_OK;
}

int cafe_b200_prune_roots(cafe_b200_ctx* c, const double* lambdas, int n_lambdas, int k, double* out)
{
    if (!c || !lambdas || !out || k < 1) return fail(c, CAFE_B200_ERR_ARG, "bad argument to prune_roots");
    CUDA_TRY(c, cudaSetDevice(c->device));
    int rc = check_counts(c);
    if (rc) return rc;
    std::vector<double> ones(c->n, 1.0), cp(k, 1.0);
    rc = stage_and_build(c, lambdas, n_lambdas, k, cp.data(), ones.data(), c->n);
    if (rc) return rc;
    double* d_root = nullptr;
    const size_t total = (size_t)c->n_families * k * c->mrf;
    CUDA_TRY(c, dev_alloc(&d_root, total));
    rc = launch_prune(c, k, CAFE_B200_GAMMA_LINSUM, d_root);
    if (rc == CAFE_B200_OK) {
        cudaError_t e = cudaEventRecord(c->ev[2], c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e == cudaSuccess) e = cudaMemcpy(out, d_root, total * sizeof(double), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(c, CAFE_B200_ERR_CUDA, cudaGetErrorString(e));
        c->ev_valid[2] = true; c->ev_valid[3] = c->ev_valid[4] = false;
    }
    cudaFree(d_root);
    return rc;
}

int cafe_b200_root_max(cafe_b200_ctx* c, const double* lambdas, int n_lambdas, double* out)
{
    if (!c || !lambdas || !out) return fail(c, CAFE_B200_ERR_ARG, "bad argument to root_max");
    CUDA_TRY(c, cudaSetDevice(c->device));
    int rc = check_counts(c);
    if (rc) return rc;
    std::vector<double> ones(c->n, 1.0);
    const double cp = 1.0;
    rc = stage_and_build(c, lambdas, n_lambdas, 1, &cp, ones.data(), c->n);
    if (rc) return rc;
    rc = launch_prune(c, 1, CAFE_B200_ROOT_MAX, nullptr);
    if (rc) return rc;
    CUDA_TRY(c, cudaEventRecord(c->ev[2], c->stream));
    c->ev_valid[2] = true; c->ev_valid[3] = c->ev_valid[4] = false;
    CUDA_TRY(c, cudaMemcpyAsync(out, c->d_cat_lk, (size_t)c->n_families * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->evals++;
    return CAFE_B200_OK;
}

int cafe_b200_pvalues(int device, const double* cond, int n_root_sizes, int n_sim, const double* observed, int64_t n_families, double* pvalues)
{
    if (!cond || !observed || !pvalues || n_root_sizes < 1 || n_sim < 1 || n_families < 0) return fail(nullptr, CAFE_B200_ERR_ARG, "bad argument to pvalues");
    if (n_sim > PV_MAX_SIM) return fail(nullptr, CAFE_B200_ERR_LIMIT, "more simulations per root size than one thread block sorts (4096)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, CAFE_B200_ERR_CUDA, "no CUDA device: this engine has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(nullptr, CAFE_B200_ERR_ARG, "device ordinal out of range");
    if (n_families == 0) return CAFE_B200_OK;
    double *d_cond = nullptr, *d_obs = nullptr, *d_p = nullptr;
    const size_t nc = (size_t)n_root_sizes * n_sim;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = dev_alloc(&d_cond, nc);
    if (e == cudaSuccess) e = dev_alloc(&d_obs, (size_t)n_families);
    if (e == cudaSuccess) e = dev_alloc(&d_p, (size_t)n_families);
    if (e == cudaSuccess) e = cudaMemcpy(d_cond, cond, nc * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_obs, observed, (size_t)n_families * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        int padded = 1;
        while (padded < n_sim) padded <<= 1;
        sort_rows_kernel<<<n_root_sizes, PV_THREADS, (size_t)padded * sizeof(double)>>>(d_cond, n_sim, padded);
        const int blocks = (int)((n_families + PV_THREADS - 1) / PV_THREADS);
        pvalue_kernel<<<blocks, PV_THREADS>>>(d_cond, n_root_sizes, n_sim, d_obs, n_families, d_p);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(pvalues, d_p, (size_t)n_families * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d_cond); cudaFree(d_obs); cudaFree(d_p);
    if (e != cudaSuccess) return fail(nullptr, CAFE_B200_ERR_CUDA, std::string("pvalues: ") + cudaGetErrorString(e));
    return CAFE_B200_OK;
}

int cafe_b200_branch_probabilities(cafe_b200_ctx* c, const double* lambdas, int n_lambdas, const int32_t* node_sizes, const uint8_t* selected,
                                   double* out)
{
    if (!c || !lambdas || !node_sizes || !out) return fail(c, CAFE_B200_ERR_ARG, "bad argument to branch_probabilities");
    if (c->n_families == 0) return CAFE_B200_OK;
    const int nn = c->tree.n_nodes;
    const size_t total = (size_t)c->n_families * nn;
    const int hi = std::min(c->mf, c->n - 1);
    for (size_t i = 0; i < total; ++i)
        if (node_sizes[i] < 0 || node_sizes[i] > hi) return fail(c, CAFE_B200_ERR_COUNT_RANGE, "a node size is outside 0..max_family_size");
    CUDA_TRY(c, cudaSetDevice(c->device));
    std::vector<double> ones(c->n, 1.0);
    const double cp = 1.0;
    int rc = stage_and_build(c, lambdas, n_lambdas, 1, &cp, ones.data(), c->n);
    if (rc) return rc;
    int32_t* d_sizes = nullptr;
    uint8_t* d_sel = nullptr;
    double* d_out = nullptr;
    cudaError_t e = dev_alloc(&d_sizes, total);
    if (e == cudaSuccess) e = dev_alloc(&d_out, total);
    if (e == cudaSuccess && selected) e = dev_alloc(&d_sel, (size_t)c->n_families);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_sizes, node_sizes, total * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && selected) e = cudaMemcpyAsync(d_sel, selected, (size_t)c->n_families, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        const int blocks = (int)((total + VT_THREADS - 1) / VT_THREADS);
        viterbi_kernel<<<blocks, VT_THREADS, 0, c->stream>>>(c->n_families, nn, c->mf, c->nr, c->d_parent, c->d_mat_of, c->d_mt, c->mt_stride, d_sizes,
                                                              d_sel, d_out);
        e = cudaGetLastError();
        c->launches++;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, total * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->stream);
    cudaFree(d_sizes); cudaFree(d_sel); cudaFree(d_out);
    if (e != cudaSuccess) return fail(c, CAFE_B200_ERR_CUDA, std::string("branch_probabilities: ") + cudaGetErrorString(e));
    return CAFE_B200_OK;
}

int cafe_b200_reconstruct(cafe_b200_ctx* c, const double* lambdas, int n_lambdas, int k, const double* prior_by_size, int32_t* states)
{
    if (!c || !lambdas || !prior_by_size || !states || k < 1) return fail(c, CAFE_B200_ERR_ARG, "bad argument to reconstruct");
    CUDA_TRY(c, cudaSetDevice(c->device));
    int rc = check_counts(c);
    if (rc) return rc;
    const int lim = std::min(c->mf, c->mrf) + 1;
    rc = stage_and_build(c, lambdas, n_lambdas, k, nullptr, prior_by_size, lim);
    if (rc) return rc;
    rc = launch_pupko(c, k, states);
    if (rc) return rc;
    c->ev_valid[2] = c->ev_valid[3] = false;
    c->ev_valid[4] = true;
    return CAFE_B200_OK;
}

int64_t cafe_b200_launch_count(const cafe_b200_ctx* c) { return c ? c->launches : 0; }

int cafe_b200_last_timings(const cafe_b200_ctx* c, double* ms4)
{
    if (!c || !ms4) return CAFE_B200_ERR_ARG;
    for (int i = 0; i < 4; ++i) ms4[i] = 0.0;
    float ms = 0.f;
    if (c->ev_valid[0] && c->ev_valid[1] && cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]) == cudaSuccess) ms4[0] = ms;
    if (c->ev_valid[1] && c->ev_valid[2] && cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]) == cudaSuccess) ms4[1] = ms;
    if (c->ev_valid[2] && c->ev_valid[3] && cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]) == cudaSuccess) ms4[2] = ms;
    if (c->ev_valid[1] && c->ev_valid[4] && cudaEventElapsedTime(&ms, c->ev[1], c->ev[4]) == cudaSuccess) ms4[3] = ms;
    cudaGetLastError();
    return CAFE_B200_OK;
}

}  // extern "C"
