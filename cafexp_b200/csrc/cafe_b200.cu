// C ABI (include/cafe_b200.h) and host orchestration of the sm_100a likelihood engine.
//
// Host work per evaluation is what the reference also does on the host before its hot loops:
// quantise (lambda, t) into matrix_cache keys (src/matrix_cache.h:47-60), derive alpha / coeff /
// log(alpha) (src/probability.cpp:150-156) — here once per unique key instead of once per matrix
// entry — and hand the device a flat program of the tree.  Everything O(families) or O(N^2) runs
// on the GPU.  There is no CPU fallback.
//
// A context owns one SHARD per CUDA device it was created on (contiguous family ranges): every
// entry point stages its host inputs once, enqueues the work on every shard's stream without waiting, then
// synchronises and combines — the score as the sum of the shards' [sum lnL, #failed] pairs in shard order on
// the host (N pinned pairs: deterministic, no collective library inside the drop-in, 16 bytes per device).
//
// Evaluation state is persistent: buffers, pinned staging, kernel attributes and the
// error-model table are set up once; an evaluation is one packed host->device copy, four kernel launches and one
// small device->host copy per shard, with no allocation and no synchronisation besides the final one.
#include "../../include/cafe_b200.h"

#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include <omp.h>
#include <thread>

#include <nvtx3/nvToolsExt.h>

#include "bd_matrix.cuh"
#include "common.cuh"
#include "plan.h"
#include "prune.cuh"
#include "prune_launch.h"
#include "pupko.cuh"
#include "pvalue.cuh"
#include "reduce.cuh"
#include "viterbi.cuh"

using namespace cafe;

namespace {

std::string g_create_error;

struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// Adds the wall time of its scope to a counter.
struct HostTimer {
    double& acc;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    explicit HostTimer(double& a) : acc(a) {}
    ~HostTimer() { acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

template <typename T>
cudaError_t dev_alloc(T** p, size_t n, bool zero = false)
{
    cudaError_t e = cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess && zero) e = cudaMemset(*p, 0, std::max<size_t>(n, 1) * sizeof(T));
    return e;
}

constexpr int FAMILY_PAD = 96;          // count rows are padded (zeros) to whole tiles of 16, 32 and 48 families
constexpr int MAX_PARTIALS = 1024;
constexpr int TIMING_HISTORY = 64;

// Offsets (bytes) of the per-evaluation parameter block: one pinned host image, one device copy per shard, moved
// with a single cudaMemcpyAsync.  Small fixed-size tables first, the pow(coeff, j) rows last, so that only the
// used prefix travels.
struct ParamLayout {
    size_t prior = 0, logprior = 0, catprobs = 0, mat_of = 0, pops = 0, leafrefs = 0, keys = 0, powc = 0, total = 0;
};

}  // namespace

// One device's share of the families and every device buffer the kernels touch.
struct Shard {
    int device = 0;
    int index = 0;
    int64_t first = 0, n_families = 0;  // family range [first, first + n_families)
    int64_t padded_families = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    int sm_count = 0;
    int smem_optin = 0;
    int max_count = 0;
    // device buffers
    void* d_counts = nullptr;           // [padded_families][n_leaves] uint8 / uint16
    void* d_raw = nullptr;              // staging for count uploads in the caller's element width
    size_t raw_bytes = 0;
    int* d_range = nullptr;             // [min, max] of an uploaded matrix
    Op* d_ops = nullptr;                // reconstruction schedule
    int* d_leaf_col = nullptr;
    int* d_parent = nullptr;
    int* d_internal = nullptr;          // [n_nodes] position among internal nodes, -1 for leaves
    unsigned char* d_param = nullptr;   // the parameter block (ParamLayout)
    // The transition matrices: one block of kstride doubles per unique key = [panelised layout | transposed layout], so that
    // a builder's slab of keys is ONE contiguous range (one peer copy / one all-gather when the build is distributed).
    double* d_mat = nullptr;
    double* d_mp = nullptr;             // = d_mat               (panelised layout of key 0)
    double* d_mt = nullptr;             // = d_mat + mp_len      (transposed layout of key 0)
    double* d_lgamma = nullptr;
    double* d_err = nullptr;
    size_t err_cap = 0;
    double* d_cat_km = nullptr;         // [k][F] category likelihoods as the pruning kernel writes them
    double* d_cat_lk = nullptr;         // [F][k] the same, family-major (finalize_kernel), what callers read
    uint8_t* d_fail = nullptr;
    double* d_family_lnl = nullptr;
    uint8_t* d_family_fail = nullptr;
    double* d_partial = nullptr;
    double* d_result = nullptr;
    double* d_scratch = nullptr;        // reconstruction spill area
    double* d_pscratch = nullptr;       // pruning: parked entries beyond tensor memory
    uint8_t* d_ctab = nullptr;          // reconstruction: argmax tables of the tiles in flight
    size_t ctab_bytes = 0;
    int32_t* d_states = nullptr;
    size_t states_cap = 0;
    double* h_result = nullptr;         // pinned [2]
    int* h_range = nullptr;             // pinned [2]
    cudaEvent_t staged = nullptr;       // H2D copy of the last parameter block has been consumed
    cudaEvent_t built = nullptr;        // this shard's slab of matrices has been built and sent to its peers
    // phase events of the last TIMING_HISTORY calls (a ring: callers that enqueue several evaluations without
    // synchronising read their durations afterwards, cafe_b200_timing_history); ev / ev_valid point at the current slot
    cudaEvent_t hist_ev[TIMING_HISTORY][5] = {};
    bool hist_valid[TIMING_HISTORY][5] = {};
    int hist_cur = 0;
    cudaEvent_t* ev = hist_ev[0];
    bool* ev_valid = hist_valid[0];
    void next_timing_slot()
    {
        hist_cur = (hist_cur + 1) % TIMING_HISTORY;
        ev = hist_ev[hist_cur];
        ev_valid = hist_valid[hist_cur];
        for (int i = 0; i < 5; ++i) ev_valid[i] = false;
    }
};

struct cafe_b200_ctx {
    std::vector<Shard*> shards;
    HostTree tree;
    int n_leaves = 0, mf = 0, mrf = 0, n = 0, mb = 0, nr = 0, kpanels = 0, n_kchunks = 0, lg_len = 0;
    int64_t n_families = 0;
    int cnt_width = 1;                  // bytes per leaf count on the device
    int max_count = 0;
    // reconstruction kernel (32-family tiles, slot machine)
    Schedule sched;
    int n_slots = 0, hw_slots = 0;
    // pruning kernel (stack machine)
    Program prog;
    PruneGeom geom = {0, 4, 2, 2, 2};
    int n_stages = 4;
    int cnt_smem_bytes = 0;
    int ops_smem_bytes = 0;             // the program of one category, staged in shared memory when it fits
    int tmem_cap = 0;                   // parked entries per warp tensor memory can hold
    int tmem_limit = -1;                // test hook: cap on tmem_cap (CAFE_B200_OPT_MAX_SLOTS)
    int tmem_entries = 0, n_gspill = 0, tmem_cols = 0, prune_smem = 0;
    int rescale = 0;
    int cap_k = 0;                      // categories the k-dependent buffers are sized for
    size_t mp_len = 0, mt_len = 0;      // doubles of one key's panelised / transposed layout
    size_t mp_stride = 0, mt_stride = 0;// doubles between consecutive keys in either layout (both = mp_len + mt_len)
    ParamLayout lay;
    // pinned host image of the parameter block and the error model
    unsigned char* h_param = nullptr;
    size_t h_param_bytes = 0;
    size_t param_used = 0;
    int n_keys = 0;                     // unique (lambda, t) keys of the staged evaluation ...
    int n_keys_padded = 0;              // ... rounded up to a whole number per builder when the build is distributed
    // Distributed matrix build: every transition matrix is built ONCE, by one device, and delivered to the others over
    // NVLink — inside a multi-device context by peer copies (below), across processes by a caller-supplied all-gather.
    bool peer_ok = false;               // every pair of this context's devices can access each other
    int ext_part = 0, ext_parts = 1;    // cafe_b200_set_build_partition
    cafe_b200_gather_fn gather_cb = nullptr;
    void* gather_user = nullptr;
    std::vector<double> err_host;       // last uploaded table
    double* h_err = nullptr;
    size_t h_err_cap = 0;
    int err_rows = 0, err_ndev = 0;
    bool has_err = false;
    int64_t launches = 0;
    int64_t evals = 0;
    // host wall time spent inside the library since create: [0] staging (keys, pow rows, program), [1] enqueueing copies
    // and kernels, [2] waiting for the devices
    double host_seconds[3] = {0.0, 0.0, 0.0};
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

cudaError_t prune_dispatch(const cafe_b200_ctx* c, const PruneParams& p, int grid, cudaStream_t s, bool attr_only)
{
    switch (c->geom.rb) {
    case 1: return prune_launch_rb1(c->geom, p, grid, c->prune_smem, s, attr_only);
    case 2: return prune_launch_rb2(c->geom, p, grid, c->prune_smem, s, attr_only);
    case 3: return prune_launch_rb3(c->geom, p, grid, c->prune_smem, s, attr_only);
    case 4: return prune_launch_rb4(c->geom, p, grid, c->prune_smem, s, attr_only);
    case 5: return prune_launch_rb5(c->geom, p, grid, c->prune_smem, s, attr_only);
    case 6: return prune_launch_rb6(c->geom, p, grid, c->prune_smem, s, attr_only);
    case 7: return prune_launch_rb7(c->geom, p, grid, c->prune_smem, s, attr_only);
    case 8: return prune_launch_rb8(c->geom, p, grid, c->prune_smem, s, attr_only);
    }
    return cudaErrorInvalidConfiguration;
}

template <int MB>
cudaError_t pupko_attr_mb(int smem) { return cudaFuncSetAttribute(pupko_kernel<MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); }

template <int MB>
int pupko_smem_mb(int slots) { return PupkoSmem<MB>::total_bytes(slots); }

#define MB_SWITCH(mb, expr)                      \
    switch (mb) {                                \
    case 1: { constexpr int MB_ = 1; expr; } break; \
    case 2: { constexpr int MB_ = 2; expr; } break; \
    case 3: { constexpr int MB_ = 3; expr; } break; \
    case 4: { constexpr int MB_ = 4; expr; } break; \
    case 5: { constexpr int MB_ = 5; expr; } break; \
    case 6: { constexpr int MB_ = 6; expr; } break; \
    case 7: { constexpr int MB_ = 7; expr; } break; \
    default: { constexpr int MB_ = 8; expr; } break; \
    }

// ------------------------------------------------------------------------------------------------------------------
// planning
// ------------------------------------------------------------------------------------------------------------------

// Shared-memory plan of the pruning kernel: the geometry with the most consumer groups whose vectors, count tile and
// at least a 3-stage matrix ring fit; the ring then takes what is left (up to MAX_RING_STAGES stages).
bool plan_prune(cafe_b200_ctx* c, int smem_limit)
{
    const int depth = c->prog.depth;
    const int cnt_tile = GFT * c->n_leaves * c->cnt_width;          // one group's count rows
    std::vector<PruneGeom> candidates;
    if (c->n <= 256) {
        const int rb = (c->n + 31) / 32;
        // three consumer groups where 160 registers per thread hold the accumulators (row blocks <= 5): 40 KB ring stages
        // (half the barrier traffic; measured 0.785 vs 0.771-0.782 of the DMMA peak with 20 KB stages), else 20 KB stages
        if (rb <= 5) candidates.push_back({rb, 4, 3, 4, 1});
        if (rb <= 5) candidates.push_back({rb, 4, 3, 2, 2});
        candidates.push_back({rb, 4, 2, 2, 2});
    }
    else {
        candidates.push_back({(c->n + 63) / 64, 8, 1, 1, 1});
    }
    if (const char* e = getenv("CAFE_B200_GEOM")) {
        // experiments: "ng,cps,pw[,stages]" for the compiled geometries of this matrix size
        int ng = 0, cps = 0, pw = 0, st = 0;
        if (sscanf(e, "%d,%d,%d,%d", &ng, &cps, &pw, &st) >= 3) {
            PruneGeom g = {candidates[0].rb, candidates[0].gw, ng, cps, pw};
            if (prune_geometry_compiled(g)) candidates.insert(candidates.begin(), g);
        }
    }
    int forced_stages = 0;
    if (const char* e = getenv("CAFE_B200_GEOM")) {
        int a, b, d, st = 0;
        if (sscanf(e, "%d,%d,%d,%d", &a, &b, &d, &st) == 4) forced_stages = st;
    }
    // the leaf list is padded to an even number of 8-byte entries so that both parts of the program copy as 16-byte words
    const int prog_bytes = (int)(c->prog.ops.size() * sizeof(POp) + ((c->prog.leaves.size() + 1) & ~size_t(1)) * sizeof(LeafRef)) + 16;
    for (const PruneGeom& g : candidates) {
        for (int variant = 0; variant < 4; ++variant) {
            // preference: counts and program on chip; then drop the program, then the counts, then both
            const bool with_counts = !(variant & 2), with_ops = !(variant & 1);
            const int cnt = with_counts ? g.ng * cnt_tile : 0;
            const int opsb = with_ops ? prog_bytes : 0;
            const int fixed = pg_total_bytes(g.rb, g.gw, g.ng, g.cps, 0, cnt, depth) + opsb;
            const int stage = pg_stage_bytes(g.rb, g.gw, g.cps);
            int stages = (smem_limit - fixed) / stage;
            stages = std::min(stages, MAX_RING_STAGES);
            if (forced_stages > 0) stages = std::min(stages, forced_stages);
            const int min_stages = 2;
            if (stages < min_stages) continue;
            c->geom = g;
            c->n_stages = stages;
            c->cnt_smem_bytes = cnt;
            c->ops_smem_bytes = opsb;
            c->prune_smem = pg_total_bytes(g.rb, g.gw, g.ng, g.cps, stages, cnt, depth) + opsb;
            c->tmem_cap = pg_tmem_capacity(g.rb, g.gw, g.ng);
            return true;
        }
    }
    return false;
}

// Parked-stack placement: the innermost entries (most frequently pushed and popped) live in tensor memory.
void place_stack(cafe_b200_ctx* c)
{
    int cap = c->tmem_cap;
    if (c->tmem_limit >= 0) cap = std::min(cap, c->tmem_limit);
    c->tmem_entries = std::min(c->prog.depth, cap);
    c->n_gspill = c->prog.depth - c->tmem_entries;
    int cols = c->geom.ng * (c->geom.gw / 4) * c->tmem_entries * pg_frag_cols(c->geom.rb);
    c->tmem_cols = 0;
    if (cols > 0) {
        c->tmem_cols = 32;
        while (c->tmem_cols < cols) c->tmem_cols <<= 1;
    }
}

int plan_pupko(cafe_b200_ctx* c, int smem_limit)
{
    if (c->mb > MAX_MB) { c->hw_slots = c->n_slots = 0; return 0; }      // reconstruction supports matrix sizes <= 256
    int base = 0, slot = 0;
    MB_SWITCH(c->mb, (base = PupkoSmem<MB_>::total_bytes(0), slot = PupkoSmem<MB_>::SLOT_BYTES));
    c->hw_slots = std::min(MAX_SLOTS, (smem_limit - base) / slot);
    c->n_slots = c->hw_slots;
    return c->hw_slots;
}

// leaf references per category in the parameter block: padded to an even count (16-byte rows)
size_t leafrefs_per_category(const cafe_b200_ctx* c) { return std::max<size_t>(2, (c->prog.leaves.size() + 1) & ~size_t(1)); }

ParamLayout make_layout(const cafe_b200_ctx* c, int k)
{
    auto up = [](size_t x) { return (x + 15) & ~size_t(15); };
    const size_t keys = (size_t)k * c->tree.n_nodes + 64;        // + padding keys of a distributed build
    ParamLayout l;
    size_t o = 0;
    l.prior = o; o = up(o + (size_t)c->n * sizeof(double));
    l.logprior = o; o = up(o + (size_t)c->n * sizeof(double));
    l.catprobs = o; o = up(o + (size_t)k * sizeof(double));
    l.mat_of = o; o = up(o + keys * sizeof(int));
    l.pops = o; o = up(o + (size_t)k * c->prog.ops.size() * sizeof(POp));
    l.leafrefs = o; o = up(o + (size_t)k * leafrefs_per_category(c) * sizeof(LeafRef));
    l.keys = o; o = up(o + keys * sizeof(KeyParams));
    l.powc = o; o = up(o + keys * c->n * sizeof(double));
    l.total = o;
    return l;
}

void free_category_buffers(Shard* s)
{
    cudaFree(s->d_param); cudaFree(s->d_mat); cudaFree(s->d_cat_lk); cudaFree(s->d_cat_km); cudaFree(s->d_fail);
    s->d_param = nullptr; s->d_mat = s->d_mp = s->d_mt = nullptr; s->d_cat_lk = s->d_cat_km = nullptr; s->d_fail = nullptr;
}

int ensure_category_buffers(cafe_b200_ctx* c, int k)
{
    if (k <= c->cap_k) return CAFE_B200_OK;
    if (k > 64) return fail(c, CAFE_B200_ERR_LIMIT, "more than 64 rate categories");
    c->cap_k = 0;
    c->lay = make_layout(c, k);
    const size_t keys = (size_t)k * c->tree.n_nodes + 64;
    for (Shard* s : c->shards) {
        CUDA_TRY(c, cudaSetDevice(s->device));
        CUDA_TRY(c, cudaStreamSynchronize(s->stream));
        free_category_buffers(s);
        CUDA_TRY(c, dev_alloc(&s->d_param, c->lay.total));
        CUDA_TRY(c, dev_alloc(&s->d_mat, keys * c->mp_stride, true));
        s->d_mp = s->d_mat;
        s->d_mt = s->d_mat + c->mp_len;
        CUDA_TRY(c, dev_alloc(&s->d_cat_lk, (size_t)s->n_families * k));
        CUDA_TRY(c, dev_alloc(&s->d_cat_km, (size_t)s->n_families * k));
        CUDA_TRY(c, dev_alloc(&s->d_fail, (size_t)s->n_families * k, true));
    }
    if (c->lay.total > c->h_param_bytes) {
        if (c->h_param) cudaFreeHost(c->h_param);
        c->h_param = nullptr;
        CUDA_TRY(c, cudaMallocHost((void**)&c->h_param, c->lay.total));
        c->h_param_bytes = c->lay.total;
    }
    c->cap_k = k;
    return CAFE_B200_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// per-evaluation staging
// ------------------------------------------------------------------------------------------------------------------

// How many builders share the matrix build of an evaluation with n_keys unique keys: the context's own devices (when
// they can reach each other and the build is big enough to outweigh the exchange), or the external partition the
// caller announced (cafe_b200_set_build_partition), else 1.
int build_partitions(const cafe_b200_ctx* c, int n_keys)
{
    if (c->ext_parts > 1) return c->gather_cb ? c->ext_parts : 1;
    const bool worth_it = (size_t)n_keys * c->n * c->n >= (size_t)4 << 20;         // >= ~0.3 ms of matrix build
    if (c->shards.size() > 1 && c->peer_ok && worth_it && !getenv("CAFE_B200_REPLICATED_BUILD")) return (int)c->shards.size();
    return 1;
}

// Quantise keys, de-duplicate, fill the pinned parameter block (once per call, shared by every shard).
int stage_host(cafe_b200_ctx* c, const double* lambdas, int n_lambdas, int k, const double* cat_probs, const double* prior, int n_prior)
{
    NvtxRange r("cafe_b200: stage keys");
    HostTimer timer(c->host_seconds[0]);
    if (n_lambdas < c->tree.n_lambdas) return fail(c, CAFE_B200_ERR_ARG, "n_lambdas smaller than the tree's lambda indices");
    for (int i = 0; i < k * n_lambdas; ++i)
        if (!std::isfinite(lambdas[i]) || lambdas[i] < 0 || lambdas[i] * 1e9 >= 9.2e18)
            return fail(c, CAFE_B200_ERR_ARG, "lambda must be finite and non-negative (lambda::is_valid, src/lambda.cpp)");
    int rc = ensure_category_buffers(c, k);
    if (rc) return rc;
    for (Shard* s : c->shards) CUDA_TRY(c, cudaEventSynchronize(s->staged));     // the previous block has left pinned memory
    const HostTree& t = c->tree;
    const size_t slots = (size_t)k * t.n_nodes;
    unsigned char* h = c->h_param;
    double* h_prior = reinterpret_cast<double*>(h + c->lay.prior);
    double* h_logprior = reinterpret_cast<double*>(h + c->lay.logprior);
    double* h_cat = reinterpret_cast<double*>(h + c->lay.catprobs);
    int* h_mat_of = reinterpret_cast<int*>(h + c->lay.mat_of);
    POp* h_pops = reinterpret_cast<POp*>(h + c->lay.pops);
    LeafRef* h_leaf = reinterpret_cast<LeafRef*>(h + c->lay.leafrefs);
    KeyParams* h_keys = reinterpret_cast<KeyParams*>(h + c->lay.keys);
    double* h_powc = reinterpret_cast<double*>(h + c->lay.powc);

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
                it = seen.emplace(std::make_pair(kl, kt), n_keys++).first;
            }
            h_mat_of[cat * t.n_nodes + v] = it->second;
        }
    }
    const int n = c->n;
    for (int j = 0; j < c->n; ++j) {
        const double pj = (prior && j < n_prior) ? prior[j] : 0.0;
        h_prior[j] = pj;
        h_logprior[j] = std::log(pj);                                        // src/base_model.cpp:98
    }
    for (int cat = 0; cat < k; ++cat) h_cat[cat] = cat_probs ? cat_probs[cat] : 1.0;
    // the pruning program with matrix slots, count columns and stack placement resolved per category
    const size_t n_ops = c->prog.ops.size(), n_leaf = c->prog.leaves.size(), leaf_row = leafrefs_per_category(c);
    for (int cat = 0; cat < k; ++cat) {
        const int* mo = h_mat_of + (size_t)cat * t.n_nodes;
        for (size_t o = 0; o < n_ops; ++o) {
            const ProgOp& po = c->prog.ops[o];
            POp q;
            q.type = po.type; q.node = po.node; q.flags = po.flags;
            q.mat = po.type == POP_GEMM ? mo[po.node] : 0;
            q.park = po.stack >= c->n_gspill ? po.stack - c->n_gspill : -(po.stack + 1);
            q.leaf_begin = po.leaf_begin; q.n_pre = po.n_pre; q.n_post = po.n_post;
            h_pops[(size_t)cat * n_ops + o] = q;
        }
        for (size_t i = 0; i < n_leaf; ++i) {
            const int leaf = c->prog.leaves[i];
            h_leaf[(size_t)cat * leaf_row + i] = {mo[leaf], t.leaf_col[leaf]};
        }
        for (size_t i = n_leaf; i < leaf_row; ++i) h_leaf[(size_t)cat * leaf_row + i] = {0, 0};
    }
    // distributed build: pad to a whole number of keys per builder with trivially "saturated" keys (one row of work)
    const int builders = build_partitions(c, n_keys);
    int padded = n_keys;
    if (builders > 1) {
        padded = (n_keys + builders - 1) / builders * builders;
        for (int key = n_keys; key < padded; ++key) {
            KeyParams kp;
            kp.saturated = 1; kp.computable = 0; kp.log_alpha = 0.0; kp.coeff = 0.0;
            h_keys[key] = kp;
            std::fill(h_powc + (size_t)key * n, h_powc + (size_t)(key + 1) * n, 0.0);
        }
    }
    // pow(coeff, j) rows from the host libm (src/probability.cpp:125): N calls per key, threaded when there are many.  A
    // context that builds only one slab of the matrices (distributed build across processes) needs only that slab's rows.
    int pow_first = 0, pow_last = n_keys;
    if (c->ext_parts > 1 && builders > 1) {
        const int per = padded / builders;
        pow_first = std::min(n_keys, c->ext_part * per);
        pow_last = std::min(n_keys, (c->ext_part + 1) * per);
    }
    const int pow_threads = std::max(1, std::min(8, omp_get_max_threads()));         // honours OMP_NUM_THREADS (torchrun sets 1 per rank)
    #pragma omp parallel for schedule(static) if ((size_t)(pow_last - pow_first) * n > 4000) num_threads(pow_threads)
    for (int key = pow_first; key < pow_last; ++key) {
        const double coeff = h_keys[key].coeff;
        double* pw = h_powc + (size_t)key * n;
        for (int j = 0; j < n; ++j) pw[j] = std::pow(coeff, (double)j);
    }
    c->n_keys = n_keys;
    c->n_keys_padded = padded;
    c->param_used = c->lay.powc + (size_t)padded * c->n * sizeof(double);
    (void)slots;
    return CAFE_B200_OK;
}

// Upload the staged block to one shard and launch the matrix builder there.
int shard_build(cafe_b200_ctx* c, Shard* s)
{
    CUDA_TRY(c, cudaSetDevice(s->device));
    cudaStream_t st = s->stream;
    s->next_timing_slot();
    CUDA_TRY(c, cudaMemcpyAsync(s->d_param, c->h_param, c->param_used, cudaMemcpyHostToDevice, st));
    CUDA_TRY(c, cudaEventRecord(s->staged, st));
    CUDA_TRY(c, cudaEventRecord(s->ev[0], st));
    // the slab of keys this shard builds
    const int builders = build_partitions(c, c->n_keys);
    const int part = c->ext_parts > 1 ? c->ext_part : s->index;
    const int per = builders > 1 ? c->n_keys_padded / builders : c->n_keys;
    const int key_first = builders > 1 ? part * per : 0;
    MatrixBuildParams mp;
    mp.n = c->n; mp.mf = c->mf; mp.nr = c->nr; mp.n_keys = c->n_keys_padded; mp.key_first = key_first; mp.lg_len = c->lg_len;
    mp.keys = reinterpret_cast<const KeyParams*>(s->d_param + c->lay.keys);
    mp.powc = reinterpret_cast<const double*>(s->d_param + c->lay.powc);
    mp.lgamma_tab = s->d_lgamma;
    mp.mp = s->d_mp; mp.mt = s->d_mt; mp.mp_stride = c->mp_stride; mp.mt_stride = c->mt_stride;
    const int entries = ((c->n + 31) / 32) * 32 * (c->mf + 1);
    int bx = (entries + MB_THREADS - 1) / MB_THREADS;
    // keep the whole launch near a few waves: many keys -> fewer blocks per key (grid-stride inside)
    const int target = std::max(1, (8 * s->sm_count + per - 1) / std::max(1, per));
    bx = std::max(1, std::min(bx, target));
    dim3 grid(bx, std::max(1, per));
    const size_t smem = ((size_t)c->lg_len + c->n) * sizeof(double);
    bd_matrix_kernel<<<grid, MB_THREADS, smem, st>>>(mp);
    CUDA_TRY(c, cudaGetLastError());
    c->launches++;
    if (builders > 1) {
        const size_t off = (size_t)key_first * c->mp_stride, cnt = (size_t)per * c->mp_stride;       // both layouts of the slab's keys
        if (c->ext_parts > 1) {
            // across processes: the caller's all-gather (NCCL on this stream) delivers every builder's slab everywhere
            const int rc = c->gather_cb(c->gather_user, s->d_mat, cnt * sizeof(double), builders, (void*)st);
            if (rc != 0) return fail(c, CAFE_B200_ERR_CUDA, "the matrix all-gather callback failed");
        }
        else {
            // inside one context: push this shard's slab to every peer over NVLink (copy engines, stream-ordered after the build)
            for (Shard* d : c->shards) {
                if (d == s) continue;
                CUDA_TRY(c, cudaMemcpyPeerAsync(d->d_mat + off, d->device, s->d_mat + off, s->device, cnt * sizeof(double), st));
            }
            CUDA_TRY(c, cudaEventRecord(s->built, st));
        }
    }
    CUDA_TRY(c, cudaEventRecord(s->ev[1], st));
    s->ev_valid[0] = s->ev_valid[1] = true;
    return CAFE_B200_OK;
}

int stage_and_build(cafe_b200_ctx* c, const double* lambdas, int n_lambdas, int k, const double* cat_probs, const double* prior, int n_prior)
{
    int rc = stage_host(c, lambdas, n_lambdas, k, cat_probs, prior, n_prior);
    if (rc) return rc;
    NvtxRange r("cafe_b200: build matrices");
    HostTimer timer(c->host_seconds[1]);
    for (Shard* s : c->shards) {
        rc = shard_build(c, s);
        if (rc) return rc;
    }
    if (c->ext_parts <= 1 && c->n_keys_padded > 0 && build_partitions(c, c->n_keys) > 1) {
        // every shard prunes only after every peer's slab has arrived in its buffers
        for (Shard* s : c->shards) {
            CUDA_TRY(c, cudaSetDevice(s->device));
            for (Shard* o : c->shards)
                if (o != s) CUDA_TRY(c, cudaStreamWaitEvent(s->stream, o->built, 0));
            CUDA_TRY(c, cudaEventRecord(s->ev[1], s->stream));       // the build phase ends when the last slab is in
        }
    }
    return CAFE_B200_OK;
}

int launch_prune(cafe_b200_ctx* c, Shard* s, int k, int mode, double* root_out)
{
    if (s->n_families == 0) return CAFE_B200_OK;
    const int pft = c->geom.ng * GFT;
    PruneParams p;
    memset(&p, 0, sizeof(p));
    p.n_families = s->n_families; p.n_leaves = c->n_leaves; p.n_nodes = c->tree.n_nodes; p.n_categories = k;
    p.mf = c->mf; p.mrf = c->mrf; p.n_ops = (int)c->prog.ops.size(); p.n_leafrefs = (int)leafrefs_per_category(c);
    p.n_kchunks = c->n_kchunks; p.mode = mode;
    p.rescale = c->rescale; p.err_rows = c->err_rows; p.err_ndev = c->err_ndev;
    p.counts_in_smem = c->cnt_smem_bytes > 0 ? 1 : 0; p.cnt_smem_bytes = c->cnt_smem_bytes; p.cnt_width = c->cnt_width;
    p.ops_in_smem = c->ops_smem_bytes > 0 ? 1 : 0;
    p.n_stages = c->n_stages; p.depth = c->prog.depth; p.n_gspill = c->n_gspill; p.tmem_entries = c->tmem_entries; p.tmem_cols = c->tmem_cols;
    p.n_tiles = (s->n_families + pft - 1) / pft;
    p.ops = reinterpret_cast<const POp*>(s->d_param + c->lay.pops);
    p.leaves = reinterpret_cast<const LeafRef*>(s->d_param + c->lay.leafrefs);
    p.counts = s->d_counts;
    p.mp = s->d_mp; p.mt = s->d_mt; p.mp_stride = c->mp_stride; p.mt_stride = c->mt_stride;
    p.err = c->has_err ? s->d_err : nullptr;
    p.prior = reinterpret_cast<const double*>(s->d_param + c->lay.prior);
    p.logprior = reinterpret_cast<const double*>(s->d_param + c->lay.logprior);
    p.cat_probs = reinterpret_cast<const double*>(s->d_param + c->lay.catprobs);
    p.scratch = s->d_pscratch;
    p.cat_lk = s->d_cat_km; p.fail = s->d_fail; p.root_out = root_out;
    const int64_t items = p.n_tiles * p.n_categories;        // one item = one tile of NG x 16 families of one category
    const int grid = (int)std::min<int64_t>(items, s->sm_count);
    CUDA_TRY(c, prune_dispatch(c, p, grid, s->stream, false));
    c->launches++;
    return CAFE_B200_OK;
}

int launch_pupko(cafe_b200_ctx* c, Shard* s, int k)
{
    if (s->n_families == 0) return CAFE_B200_OK;
    const HostTree& t = c->tree;
    const int64_t n_tiles = (s->n_families + FT - 1) / FT;
    const int grid = (int)std::min<int64_t>(n_tiles * k, s->sm_count);
    const size_t n_states = (size_t)s->n_families * k * t.n_internal;
    const size_t ctab = (size_t)grid * t.n_internal * FT * c->nr;
    if (ctab > s->ctab_bytes) {
        CUDA_TRY(c, cudaStreamSynchronize(s->stream));
        cudaFree(s->d_ctab); s->d_ctab = nullptr; s->ctab_bytes = 0;
        CUDA_TRY(c, dev_alloc(&s->d_ctab, ctab));
        s->ctab_bytes = ctab;
    }
    if (n_states > s->states_cap) {
        CUDA_TRY(c, cudaStreamSynchronize(s->stream));
        cudaFree(s->d_states); s->d_states = nullptr; s->states_cap = 0;
        CUDA_TRY(c, dev_alloc(&s->d_states, n_states));
        s->states_cap = n_states;
    }
    PupkoParams p;
    memset(&p, 0, sizeof(p));
    p.n_families = s->n_families; p.n_leaves = c->n_leaves; p.n_nodes = t.n_nodes; p.n_internal = t.n_internal; p.n_categories = k;
    p.mf = c->mf; p.mrf = c->mrf; p.n_ops = (int)c->sched.ops.size(); p.n_kchunks = c->n_kchunks;
    p.n_spill = std::max(1, c->sched.n_spill); p.n_slots = c->n_slots;
    p.counts_in_smem = (FT * c->n_leaves * 2 <= CNT_CAP_BYTES) ? 1 : 0; p.cnt_width = c->cnt_width;
    p.n_tiles = n_tiles; p.ops = s->d_ops; p.counts = s->d_counts; p.leaf_col = s->d_leaf_col; p.parent = s->d_parent;
    p.internal_idx = s->d_internal; p.mat_of = reinterpret_cast<const int*>(s->d_param + c->lay.mat_of);
    p.mt = s->d_mt; p.mt_stride = c->mt_stride; p.prior = reinterpret_cast<const double*>(s->d_param + c->lay.prior);
    p.scratch = s->d_scratch; p.ctab = s->d_ctab; p.states = s->d_states;
    int smem = 0;
    MB_SWITCH(c->mb, (smem = PupkoSmem<MB_>::total_bytes(c->n_slots), pupko_kernel<MB_><<<grid, pupko_threads(PUPKO_WARPS), smem, s->stream>>>(p)));
    CUDA_TRY(c, cudaGetLastError());
    c->launches++;
    return CAFE_B200_OK;
}

// (Re)build the reconstruction schedule for n_slots slots and size the spill scratch of both kernels on every shard.
int upload_plans(cafe_b200_ctx* c)
{
    if (c->n_slots >= 2) c->sched = ScheduleBuilder(c->tree, c->n_slots).build();
    place_stack(c);
    const size_t ldv = ldv_of(std::min(c->mb, MAX_MB));
    const size_t consumer_threads = (size_t)c->geom.ng * c->geom.gw * 32;
    for (Shard* s : c->shards) {
        CUDA_TRY(c, cudaSetDevice(s->device));
        CUDA_TRY(c, cudaStreamSynchronize(s->stream));
        cudaFree(s->d_ops); cudaFree(s->d_scratch); cudaFree(s->d_pscratch);
        s->d_ops = nullptr; s->d_scratch = nullptr; s->d_pscratch = nullptr;
        if (c->n_slots >= 2) {
            CUDA_TRY(c, dev_alloc(&s->d_ops, c->sched.ops.size()));
            CUDA_TRY(c, cudaMemcpy(s->d_ops, c->sched.ops.data(), c->sched.ops.size() * sizeof(Op), cudaMemcpyHostToDevice));
            CUDA_TRY(c, dev_alloc(&s->d_scratch, (size_t)s->sm_count * std::max(1, c->sched.n_spill) * FT * ldv, true));
        }
        CUDA_TRY(c, dev_alloc(&s->d_pscratch, (size_t)s->sm_count * std::max(1, c->n_gspill) * consumer_threads * 4 * c->geom.rb, true));
    }
    return CAFE_B200_OK;
}

int check_counts(cafe_b200_ctx* c)
{
    int hi = c->max_count;
    if (c->has_err) {
        if (c->max_count >= c->err_rows) return fail(c, CAFE_B200_ERR_COUNT_RANGE, "a leaf count has no error-model row");
        hi += (c->err_ndev - 1) / 2;
    }
    if (hi > c->mf) return fail(c, CAFE_B200_ERR_COUNT_RANGE, "a leaf count (plus error-model deviation) exceeds max_family_size");
    return CAFE_B200_OK;
}

// Upload a count matrix in the caller's element width; narrow it to the device width and range-check it on the device.
int upload_counts(cafe_b200_ctx* c, const void* counts, int count_bytes)
{
    NvtxRange r("cafe_b200: upload families");
    if (count_bytes != 1 && count_bytes != 2 && count_bytes != 4) return fail(c, CAFE_B200_ERR_ARG, "count_bytes must be 1, 2 or 4");
    const size_t row = (size_t)c->n_leaves;
    for (Shard* s : c->shards) {
        if (s->n_families == 0) continue;
        CUDA_TRY(c, cudaSetDevice(s->device));
        const size_t total = (size_t)s->n_families * row;
        const unsigned char* src = static_cast<const unsigned char*>(counts) + (size_t)s->first * row * count_bytes;
        s->h_range[0] = INT_MAX; s->h_range[1] = INT_MIN;
        CUDA_TRY(c, cudaMemcpyAsync(s->d_range, s->h_range, 2 * sizeof(int), cudaMemcpyHostToDevice, s->stream));
        const void* d_src = s->d_counts;
        if (count_bytes == c->cnt_width) {
            CUDA_TRY(c, cudaMemcpyAsync(s->d_counts, src, total * count_bytes, cudaMemcpyHostToDevice, s->stream));
        }
        else {
            if (total * count_bytes > s->raw_bytes) {
                CUDA_TRY(c, cudaStreamSynchronize(s->stream));
                cudaFree(s->d_raw); s->d_raw = nullptr; s->raw_bytes = 0;
                CUDA_TRY(c, cudaMalloc(&s->d_raw, total * count_bytes));
                s->raw_bytes = total * count_bytes;
            }
            CUDA_TRY(c, cudaMemcpyAsync(s->d_raw, src, total * count_bytes, cudaMemcpyHostToDevice, s->stream));
            d_src = s->d_raw;
        }
        const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(((int64_t)total / 4 + RED_THREADS - 1) / RED_THREADS, 8 * (int64_t)s->sm_count));
        ingest_counts_kernel<<<blocks, RED_THREADS, 0, s->stream>>>(d_src, count_bytes, s->d_counts, c->cnt_width, (int64_t)total, c->mf, s->d_range);
        CUDA_TRY(c, cudaGetLastError());
        c->launches++;
        CUDA_TRY(c, cudaMemcpyAsync(s->h_range, s->d_range, 2 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    }
    int lo = INT_MAX, hi = 0;
    for (Shard* s : c->shards) {
        if (s->n_families == 0) continue;
        CUDA_TRY(c, cudaSetDevice(s->device));
        CUDA_TRY(c, cudaStreamSynchronize(s->stream));
        lo = std::min(lo, s->h_range[0]);
        hi = std::max(hi, s->h_range[1]);
    }
    // the context now holds the new matrix; out-of-range counts are refused here and again by every evaluation
    c->max_count = hi;
    if (lo < 0) { c->max_count = c->mf + 1; return fail(c, CAFE_B200_ERR_COUNT_RANGE, "negative leaf count"); }
    if (hi > c->mf) return fail(c, CAFE_B200_ERR_COUNT_RANGE, "a leaf count exceeds max_family_size");
    return CAFE_B200_OK;
}

int enqueue_eval(cafe_b200_ctx* c, const double* lambdas, int n_lambdas, const double* cat_probs, int k, const double* prior, int mode,
                 double* result_device_single)
{
    if (!c || !lambdas || k < 1 || n_lambdas < 1 || (mode != CAFE_B200_BASE_LOGMAX && mode != CAFE_B200_GAMMA_LINSUM))
        return fail(c, CAFE_B200_ERR_ARG, "bad argument to eval");
    if (mode == CAFE_B200_BASE_LOGMAX && k != 1) return fail(c, CAFE_B200_ERR_ARG, "base mode takes exactly one category");
    if (!prior) return fail(c, CAFE_B200_ERR_ARG, "prior is required");
    int rc = check_counts(c);
    if (rc) return rc;
    rc = stage_and_build(c, lambdas, n_lambdas, k, cat_probs, prior, c->mrf);
    if (rc) return rc;
    NvtxRange r("cafe_b200: prune + reduce");
    HostTimer timer(c->host_seconds[1]);
    for (Shard* s : c->shards) {
        CUDA_TRY(c, cudaSetDevice(s->device));
        rc = launch_prune(c, s, k, mode, nullptr);
        if (rc) return rc;
        cudaStream_t st = s->stream;
        CUDA_TRY(c, cudaEventRecord(s->ev[2], st));
        const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((s->n_families + RED_THREADS - 1) / RED_THREADS, MAX_PARTIALS));
        finalize_kernel<<<blocks, RED_THREADS, 0, st>>>(s->n_families, k, mode, s->d_cat_km, s->d_fail, s->d_family_lnl, s->d_family_fail, s->d_partial, s->d_cat_lk);
        CUDA_TRY(c, cudaGetLastError());
        final_sum_kernel<<<1, RED_THREADS, 0, st>>>(blocks, s->d_partial, result_device_single ? result_device_single : s->d_result);
        CUDA_TRY(c, cudaGetLastError());
        c->launches += 2;
        CUDA_TRY(c, cudaEventRecord(s->ev[3], st));
        s->ev_valid[2] = s->ev_valid[3] = true;
        s->ev_valid[4] = false;
    }
    c->evals++;
    return CAFE_B200_OK;
}

int sync_all(cafe_b200_ctx* c)
{
    HostTimer timer(c->host_seconds[2]);
    for (Shard* s : c->shards) {
        CUDA_TRY(c, cudaSetDevice(s->device));
        CUDA_TRY(c, cudaStreamSynchronize(s->stream));
    }
    return CAFE_B200_OK;
}

void destroy_shard(Shard* s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    cudaFree(s->d_counts); cudaFree(s->d_raw); cudaFree(s->d_range); cudaFree(s->d_ops); cudaFree(s->d_leaf_col); cudaFree(s->d_parent);
    cudaFree(s->d_internal); cudaFree(s->d_lgamma); cudaFree(s->d_err); cudaFree(s->d_family_lnl); cudaFree(s->d_family_fail);
    cudaFree(s->d_partial); cudaFree(s->d_result); cudaFree(s->d_scratch); cudaFree(s->d_pscratch); cudaFree(s->d_ctab); cudaFree(s->d_states);
    free_category_buffers(s);
    if (s->h_result) cudaFreeHost(s->h_result);
    if (s->h_range) cudaFreeHost(s->h_range);
    if (s->staged) cudaEventDestroy(s->staged);
    if (s->built) cudaEventDestroy(s->built);
    for (auto& slot : s->hist_ev)
        for (auto& e : slot) if (e) cudaEventDestroy(e);
    if (s->own_stream) cudaStreamDestroy(s->own_stream);
    delete s;
}

}  // namespace

extern "C" {

int cafe_b200_abi_version(void) { return CAFE_B200_ABI_VERSION; }

void cafe_b200_get_limits(cafe_b200_limits* out)
{
    if (!out) return;
    out->max_matrix_size = 512;                 // likelihood evaluation; reconstruction: 32 * MAX_MB = 256
    out->max_categories = 64;
    out->max_nodes = 1 << 20;
    out->families_per_tile = MAX_GROUPS * GFT;
}

int cafe_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* cafe_b200_last_error(const cafe_b200_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

void* cafe_b200_alloc_pinned(size_t bytes)
{
    void* p = nullptr;
    if (cudaMallocHost(&p, std::max<size_t>(bytes, 1)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void cafe_b200_free_pinned(void* p)
{
    if (p) cudaFreeHost(p);
}

void cafe_b200_destroy(cafe_b200_ctx* c)
{
    if (!c) return;
    for (Shard* s : c->shards) destroy_shard(s);
    if (c->h_param) cudaFreeHost(c->h_param);
    if (c->h_err) cudaFreeHost(c->h_err);
    delete c;
}

int cafe_b200_create_multi(cafe_b200_ctx** out, const cafe_b200_tree* tree, const void* leaf_counts, int count_bytes, int64_t n_families,
                           int n_leaves, int max_family_size, int max_root_family_size, const int* devices, int n_devices)
{
    if (!out || !tree || n_families < 0 || n_leaves < 1 || max_family_size < 1 || max_root_family_size < 1 || (n_families > 0 && !leaf_counts) ||
        !devices || n_devices < 1)
        return fail(nullptr, CAFE_B200_ERR_ARG, "bad argument to create");
    *out = nullptr;
    const int nn = tree->n_nodes;
    if (nn < 2 || !tree->parent || !tree->child_offset || !tree->child_list || !tree->leaf_col || !tree->branch || !tree->lambda_index)
        return fail(nullptr, CAFE_B200_ERR_ARG, "incomplete tree");
    if (count_bytes != 1 && count_bytes != 2 && count_bytes != 4) return fail(nullptr, CAFE_B200_ERR_ARG, "count_bytes must be 1, 2 or 4");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, CAFE_B200_ERR_CUDA, "no CUDA device: this engine has no CPU fallback");
    }
    for (int i = 0; i < n_devices; ++i) {
        if (devices[i] < 0 || devices[i] >= ndev) return fail(nullptr, CAFE_B200_ERR_ARG, "device ordinal out of range");
        for (int j = 0; j < i; ++j)
            if (devices[j] == devices[i]) return fail(nullptr, CAFE_B200_ERR_ARG, "a device is listed twice");
    }

    cafe_b200_ctx* c = new cafe_b200_ctx();
    HostTree& t = c->tree;
    {
        const char* why = import_tree(t, tree, n_leaves);
        if (why) { g_create_error = why; delete c; return CAFE_B200_ERR_ARG; }
    }
    c->n_leaves = n_leaves; c->mf = max_family_size; c->mrf = max_root_family_size;
    c->n = std::max(c->mf, c->mrf) + 1;                                      // src/base_model.cpp:77
    c->n_families = n_families;
    if (c->n > 512) {
        g_create_error = "matrix size beyond this build (max 512)";
        delete c;
        return CAFE_B200_ERR_LIMIT;
    }
    c->cnt_width = c->mf <= 255 ? 1 : 2;
    c->lg_len = 2 * c->n + 2;                    // lgamma arguments reach s + c <= 2 (N - 1)           src/probability.cpp:58-64
    c->prog = ProgramBuilder(t).build();

#define CREATE_TRY(call)                                                                      \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            g_create_error = std::string(#call) + ": " + cudaGetErrorString(e__);             \
            cafe_b200_destroy(c);                                                             \
            return CAFE_B200_ERR_CUDA;                                                        \
        }                                                                                     \
    } while (0)

    if (n_devices > 1) {
        // primary contexts are created concurrently (0.3 s each when done one after the other)
        std::vector<std::thread> warm;
        for (int i = 0; i < n_devices; ++i) warm.emplace_back([dev = devices[i]] { if (cudaSetDevice(dev) == cudaSuccess) cudaFree(nullptr); });
        for (std::thread& t : warm) t.join();
        cudaGetLastError();
    }
    int smem_limit = INT_MAX;
    for (int i = 0; i < n_devices; ++i) {
        Shard* s = new Shard();
        c->shards.push_back(s);
        s->device = devices[i];
        s->index = i;
        s->first = n_families * i / n_devices;
        s->n_families = n_families * (i + 1) / n_devices - s->first;
        s->padded_families = (s->n_families + FAMILY_PAD - 1) / FAMILY_PAD * FAMILY_PAD + FAMILY_PAD;
        CREATE_TRY(cudaSetDevice(s->device));
        CREATE_TRY(cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, s->device));
        CREATE_TRY(cudaDeviceGetAttribute(&s->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, s->device));
        smem_limit = std::min(smem_limit, s->smem_optin);
        CREATE_TRY(cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking));
        s->stream = s->own_stream;
        CREATE_TRY(cudaEventCreateWithFlags(&s->staged, cudaEventDisableTiming));
        CREATE_TRY(cudaEventCreateWithFlags(&s->built, cudaEventDisableTiming));
        CREATE_TRY(cudaEventRecord(s->staged, s->stream));
        for (auto& slot : s->hist_ev)
            for (auto& e : slot) CREATE_TRY(cudaEventCreate(&e));
    }

    if (!plan_prune(c, smem_limit)) { g_create_error = "not enough shared memory for the pruning kernel at this matrix size / tree"; cafe_b200_destroy(c); return CAFE_B200_ERR_LIMIT; }
    c->nr = pg_nr(c->geom.rb, c->geom.gw);
    c->mb = c->nr / 32;
    c->kpanels = ((c->mf + 1 + 3) / 4 + PPS - 1) / PPS * PPS;
    c->n_kchunks = c->kpanels / PPS;
    c->mp_len = (size_t)c->kpanels * c->nr * 4;
    c->mt_len = (size_t)c->kpanels * 4 * c->nr;         // columns padded to whole ring stages (zeros)
    c->mp_stride = c->mt_stride = c->mp_len + c->mt_len;
    plan_pupko(c, smem_limit);

    if (c->shards.size() > 1) {
        c->peer_ok = true;
        for (Shard* a : c->shards)
            for (Shard* b : c->shards) {
                if (a == b) continue;
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, a->device, b->device) != cudaSuccess || !can) { c->peer_ok = false; continue; }
                cudaSetDevice(a->device);
                const cudaError_t e = cudaDeviceEnablePeerAccess(b->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) c->peer_ok = false;
                cudaGetLastError();
            }
    }
    std::vector<double> lg(c->lg_len);
    for (int i = 0; i < c->lg_len; ++i) lg[i] = lgamma((double)i);           // src/probability.cpp:58-72 (table below 1024, libm above: same values)
    std::vector<int> internal_idx(nn, -1);
    {
        int ni = 0;
        for (int v = 0; v < nn; ++v) if (!t.is_leaf(v)) internal_idx[v] = ni++;
    }
    for (Shard* s : c->shards) {
        CREATE_TRY(cudaSetDevice(s->device));
        {
            PruneParams dummy;
            memset(&dummy, 0, sizeof(dummy));
            CREATE_TRY(prune_dispatch(c, dummy, 1, s->stream, true));
            if (c->hw_slots >= 2) {
                int smem = 0;
                cudaError_t e = cudaSuccess;
                MB_SWITCH(c->mb, (smem = PupkoSmem<MB_>::total_bytes(c->hw_slots), e = pupko_attr_mb<MB_>(smem)));
                CREATE_TRY(e);
            }
        }
        CREATE_TRY(cudaMalloc(&s->d_counts, (size_t)s->padded_families * n_leaves * c->cnt_width));
        CREATE_TRY(cudaMemsetAsync(s->d_counts, 0, (size_t)s->padded_families * n_leaves * c->cnt_width, s->stream));
        CREATE_TRY(dev_alloc(&s->d_range, (size_t)2));
        CREATE_TRY(cudaMallocHost((void**)&s->h_range, 2 * sizeof(int)));
        CREATE_TRY(dev_alloc(&s->d_leaf_col, (size_t)nn));
        CREATE_TRY(cudaMemcpy(s->d_leaf_col, t.leaf_col.data(), nn * sizeof(int), cudaMemcpyHostToDevice));
        CREATE_TRY(dev_alloc(&s->d_parent, (size_t)nn));
        CREATE_TRY(cudaMemcpy(s->d_parent, t.parent.data(), nn * sizeof(int), cudaMemcpyHostToDevice));
        CREATE_TRY(dev_alloc(&s->d_internal, (size_t)nn));
        CREATE_TRY(cudaMemcpy(s->d_internal, internal_idx.data(), nn * sizeof(int), cudaMemcpyHostToDevice));
        CREATE_TRY(dev_alloc(&s->d_lgamma, (size_t)c->lg_len));
        CREATE_TRY(cudaMemcpy(s->d_lgamma, lg.data(), c->lg_len * sizeof(double), cudaMemcpyHostToDevice));
        CREATE_TRY(dev_alloc(&s->d_family_lnl, (size_t)s->n_families));
        CREATE_TRY(dev_alloc(&s->d_family_fail, (size_t)s->n_families, true));
        CREATE_TRY(dev_alloc(&s->d_partial, (size_t)2 * MAX_PARTIALS));
        CREATE_TRY(dev_alloc(&s->d_result, (size_t)2));
        CREATE_TRY(cudaMallocHost((void**)&s->h_result, 2 * sizeof(double)));
    }
#undef CREATE_TRY
    if (upload_plans(c) != CAFE_B200_OK) { g_create_error = c->error; cafe_b200_destroy(c); return CAFE_B200_ERR_CUDA; }
    if (n_families > 0) {
        int rc = upload_counts(c, leaf_counts, count_bytes);
        if (rc) { g_create_error = c->error; cafe_b200_destroy(c); return rc; }
    }
    *out = c;
    return CAFE_B200_OK;
}

int cafe_b200_create(cafe_b200_ctx** out, const cafe_b200_tree* tree, const int32_t* leaf_counts, int64_t n_families, int n_leaves,
                     int max_family_size, int max_root_family_size, int device)
{
    return cafe_b200_create_multi(out, tree, leaf_counts, 4, n_families, n_leaves, max_family_size, max_root_family_size, &device, 1);
}

int cafe_b200_n_devices(const cafe_b200_ctx* c) { return c ? (int)c->shards.size() : 0; }

int cafe_b200_set_families_ex(cafe_b200_ctx* c, const void* leaf_counts, int count_bytes, int64_t n_families)
{
    if (!c || !leaf_counts || n_families != c->n_families) return fail(c, CAFE_B200_ERR_ARG, "set_families: shape must match create");
    if (n_families * c->n_leaves == 0) return CAFE_B200_OK;
    return upload_counts(c, leaf_counts, count_bytes);
}

int cafe_b200_set_families(cafe_b200_ctx* c, const int32_t* leaf_counts, int64_t n_families)
{
    return cafe_b200_set_families_ex(c, leaf_counts, 4, n_families);
}

int cafe_b200_set_error_model(cafe_b200_ctx* c, const double* probs, int rows, int n_deviations)
{
    if (!c) return CAFE_B200_ERR_ARG;
    if (!probs) { c->has_err = false; c->err_rows = c->err_ndev = 0; c->err_host.clear(); return CAFE_B200_OK; }
    if (rows < 1 || n_deviations < 1 || (n_deviations % 2) == 0) return fail(c, CAFE_B200_ERR_ARG, "error model: odd number of deviations required");
    const size_t n = (size_t)rows * n_deviations;
    // the epsilon optimiser calls this before every evaluation: nothing to do unless the table changed
    if (c->has_err && rows == c->err_rows && n_deviations == c->err_ndev && c->err_host.size() == n &&
        memcmp(c->err_host.data(), probs, n * sizeof(double)) == 0)
        return CAFE_B200_OK;
    if (n > c->h_err_cap) {
        for (Shard* s : c->shards) { CUDA_TRY(c, cudaSetDevice(s->device)); CUDA_TRY(c, cudaStreamSynchronize(s->stream)); }
        if (c->h_err) cudaFreeHost(c->h_err);
        c->h_err = nullptr; c->h_err_cap = 0;
        CUDA_TRY(c, cudaMallocHost((void**)&c->h_err, n * sizeof(double)));
        c->h_err_cap = n;
    }
    else {
        for (Shard* s : c->shards) CUDA_TRY(c, cudaEventSynchronize(s->staged));      // the previous table has left pinned memory
    }
    memcpy(c->h_err, probs, n * sizeof(double));
    c->err_host.assign(probs, probs + n);
    for (Shard* s : c->shards) {
        CUDA_TRY(c, cudaSetDevice(s->device));
        if (n > s->err_cap) {
            CUDA_TRY(c, cudaStreamSynchronize(s->stream));
            cudaFree(s->d_err); s->d_err = nullptr; s->err_cap = 0;
            CUDA_TRY(c, dev_alloc(&s->d_err, n));
            s->err_cap = n;
        }
        CUDA_TRY(c, cudaMemcpyAsync(s->d_err, c->h_err, n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(c, cudaEventRecord(s->staged, s->stream));
    }
    c->err_rows = rows; c->err_ndev = n_deviations; c->has_err = true;
    return CAFE_B200_OK;
}

int cafe_b200_set_option(cafe_b200_ctx* c, int option, int value)
{
    if (!c) return CAFE_B200_ERR_ARG;
    if (option == CAFE_B200_OPT_RESCALE) { c->rescale = value ? 1 : 0; return CAFE_B200_OK; }
    if (option == CAFE_B200_OPT_MAX_SLOTS) {
        if (value < 2) return fail(c, CAFE_B200_ERR_ARG, "at least two slots are required");
        c->n_slots = std::min(value, c->hw_slots);
        c->tmem_limit = value - 2;               // pruning: value - 2 parked entries may use tensor memory, the rest spill
        return upload_plans(c);
    }
    return fail(c, CAFE_B200_ERR_ARG, "unknown option");
}

// Flat reader of the CAFE tab format: what read_gene_families (src/io.cpp:134-215, "CAFE input
// format" branch) + gene_family::get_species_size produce, without a std::map per family in between.
int cafe_b200_read_family_table(const char* path, const char* const* leaf_names, int n_leaves, int32_t* counts, int64_t cap_families,
                                int64_t* n_families, char* ids, int id_stride)
{
    if (!path || !leaf_names || n_leaves < 1 || !n_families) return fail(nullptr, CAFE_B200_ERR_ARG, "bad argument to read_family_table");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(nullptr, CAFE_B200_ERR_ARG, std::string("cannot open ") + path);
    std::string text;
    {
        char buf[1 << 16];
        size_t got;
        while ((got = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, got);
        fclose(f);
    }
    auto lower = [](std::string s) { for (char& ch : s) ch = (char)tolower((unsigned char)ch); return s; };
    std::map<std::string, int> col_of;                         // species name (lower case, as gene_family keys compare) -> count column
    for (int l = 0; l < n_leaves; ++l) col_of[lower(leaf_names[l])] = l;
    size_t pos = 0;
    auto next_line = [&](size_t& b, size_t& e) {
        if (pos >= text.size()) return false;
        b = pos;
        size_t nl = text.find('\n', pos);
        e = nl == std::string::npos ? text.size() : nl;
        pos = e + 1;
        if (e > b && text[e - 1] == '\r') --e;
        return true;
    };
    size_t b = 0, e = 0;
    if (!next_line(b, e)) return fail(nullptr, CAFE_B200_ERR_ARG, "No families found");
    if (text[b] == '#') return fail(nullptr, CAFE_B200_ERR_ARG, "the '#'-header (CAFExp) family format is not supported by the flat reader");
    // header: Desc <TAB> Family ID <TAB> species ...      (src/io.cpp:160-170: every column from the third on is a species)
    std::vector<int> col_of_field;
    {
        size_t s = b;
        int field = 0;
        std::vector<bool> seen(n_leaves, false);
        while (s <= e) {
            size_t t = text.find('\t', s);
            if (t == std::string::npos || t > e) t = e;
            if (field >= 2) {
                auto it = col_of.find(lower(text.substr(s, t - s)));
                col_of_field.push_back(it == col_of.end() ? -1 : it->second);
                if (it != col_of.end()) seen[it->second] = true;
            }
            ++field;
            s = t + 1;
        }
        for (int l = 0; l < n_leaves; ++l)
            if (!seen[l]) return fail(nullptr, CAFE_B200_ERR_ARG, std::string("species missing from the family table: ") + leaf_names[l]);
    }
    int64_t n = 0;
    while (next_line(b, e)) {
        if (e == b) continue;
        // fields: description, id, counts (atoi semantics, src/io.cpp:186)
        size_t s = b;
        int field = 0;
        bool row_ok = n < cap_families && counts;
        if (row_ok) std::fill(counts + n * n_leaves, counts + (n + 1) * n_leaves, 0);
        while (s <= e) {
            size_t t = text.find('\t', s);
            if (t == std::string::npos || t > e) t = e;
            if (field == 1 && ids && n < cap_families && id_stride > 0) {
                const size_t len = std::min<size_t>(t - s, (size_t)id_stride - 1);
                memcpy(ids + n * id_stride, text.data() + s, len);
                ids[n * id_stride + len] = 0;
            }
            else if (field >= 2 && row_ok && (size_t)(field - 2) < col_of_field.size()) {
                const int col = col_of_field[field - 2];
                if (col >= 0) counts[n * n_leaves + col] = atoi(text.c_str() + s);
            }
            ++field;
            s = t + 1;
        }
        if (field < 3) continue;                                 // not a family line
        ++n;
    }
    *n_families = n;
    if (n == 0) return fail(nullptr, CAFE_B200_ERR_ARG, "No families found");
    return CAFE_B200_OK;
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

int cafe_b200_plan_program(const cafe_b200_tree* tree, int* ops_out, int cap, int* n_ops, int* leaves_out, int leaves_cap, int* n_leaf_refs,
                           int* depth)
{
    if (!tree || tree->n_nodes < 2 || !n_ops) return CAFE_B200_ERR_ARG;
    HostTree t;
    const char* why = import_tree(t, tree, -1);
    if (why) { g_create_error = why; return CAFE_B200_ERR_ARG; }
    Program p = ProgramBuilder(t).build();
    *n_ops = (int)p.ops.size();
    if (n_leaf_refs) *n_leaf_refs = (int)p.leaves.size();
    if (depth) *depth = p.depth;
    if (ops_out) {
        if (cap < (int)p.ops.size()) return CAFE_B200_ERR_ARG;
        for (size_t i = 0; i < p.ops.size(); ++i) {
            const ProgOp& o = p.ops[i];
            int* q = ops_out + 7 * i;
            q[0] = o.type; q[1] = o.node; q[2] = o.flags; q[3] = o.stack; q[4] = o.leaf_begin; q[5] = o.n_pre; q[6] = o.n_post;
        }
    }
    if (leaves_out) {
        if (leaves_cap < (int)p.leaves.size()) return CAFE_B200_ERR_ARG;
        std::copy(p.leaves.begin(), p.leaves.end(), leaves_out);
    }
    return CAFE_B200_OK;
}

int cafe_b200_set_build_partition(cafe_b200_ctx* c, int part, int n_parts, cafe_b200_gather_fn gather, void* user)
{
    if (!c) return CAFE_B200_ERR_ARG;
    if (c->shards.size() != 1) return fail(c, CAFE_B200_ERR_ARG, "set_build_partition: single-device contexts only (a multi-device context distributes by itself)");
    if (n_parts < 1 || part < 0 || part >= n_parts || n_parts > 64 || (n_parts > 1 && !gather)) return fail(c, CAFE_B200_ERR_ARG, "bad build partition");
    c->ext_part = part; c->ext_parts = n_parts; c->gather_cb = gather; c->gather_user = user;
    return CAFE_B200_OK;
}

int cafe_b200_set_stream(cafe_b200_ctx* c, void* cuda_stream)
{
    if (!c) return CAFE_B200_ERR_ARG;
    if (c->shards.size() != 1) return fail(c, CAFE_B200_ERR_ARG, "set_stream: the context spans several devices");
    Shard* s = c->shards[0];
    CUDA_TRY(c, cudaSetDevice(s->device));
    CUDA_TRY(c, cudaStreamSynchronize(s->stream));
    s->stream = (cudaStream_t)cuda_stream;          // NULL is the legacy default stream, as in the CUDA runtime
    CUDA_TRY(c, cudaEventRecord(s->staged, s->stream));
    return CAFE_B200_OK;
}

int cafe_b200_eval_device(cafe_b200_ctx* c, const double* lambdas, int n_lambdas, const double* cat_probs, int k, const double* prior,
                          int mode, double* result_device)
{
    if (!c) return CAFE_B200_ERR_ARG;
    if (!result_device) return fail(c, CAFE_B200_ERR_ARG, "result_device is null");
    if (c->shards.size() != 1) return fail(c, CAFE_B200_ERR_ARG, "eval_device: the context spans several devices; use cafe_b200_eval");
    return enqueue_eval(c, lambdas, n_lambdas, cat_probs, k, prior, mode, result_device);
}

int cafe_b200_eval(cafe_b200_ctx* c, const double* lambdas, int n_lambdas, const double* cat_probs, int k, const double* prior, int mode,
                   double* neg_lnl, double* family_lnl, double* cat_lk, int64_t* n_failed, int64_t* failed_idx, int64_t failed_cap)
{
    if (!c || !neg_lnl) return fail(c, CAFE_B200_ERR_ARG, "neg_lnl is null");
    NvtxRange r("cafe_b200_eval");
    int rc = enqueue_eval(c, lambdas, n_lambdas, cat_probs, k, prior, mode, nullptr);
    if (rc) return rc;
    for (Shard* s : c->shards) {
        CUDA_TRY(c, cudaSetDevice(s->device));
        cudaStream_t st = s->stream;
        CUDA_TRY(c, cudaMemcpyAsync(s->h_result, s->d_result, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (family_lnl && s->n_families)
            CUDA_TRY(c, cudaMemcpyAsync(family_lnl + s->first, s->d_family_lnl, (size_t)s->n_families * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (cat_lk && s->n_families && mode == CAFE_B200_GAMMA_LINSUM)
            CUDA_TRY(c, cudaMemcpyAsync(cat_lk + (size_t)s->first * k, s->d_cat_lk, (size_t)s->n_families * k * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    rc = sync_all(c);
    if (rc) return rc;
    // the shards' pairs are added on the host in shard order: deterministic for a given device list
    double sum = 0.0;
    int64_t nf = 0;
    for (Shard* s : c->shards) {
        if (s->n_families == 0) continue;
        sum += s->h_result[0];
        nf += (int64_t)s->h_result[1];
    }
    if (n_failed) *n_failed = nf;
    *neg_lnl = nf > 0 ? INFINITY : -sum;                                  // src/gamma_core.cpp:227-236 / src/base_model.cpp:107
    if (nf > 0 && failed_idx && failed_cap > 0) {
        int64_t w = 0;
        for (Shard* s : c->shards) {
            if (s->n_families == 0 || w >= failed_cap) continue;
            std::vector<uint8_t> flags((size_t)s->n_families);
            CUDA_TRY(c, cudaSetDevice(s->device));
            CUDA_TRY(c, cudaMemcpy(flags.data(), s->d_family_fail, flags.size(), cudaMemcpyDeviceToHost));
            for (int64_t i = 0; i < s->n_families && w < failed_cap; ++i)
                if (flags[i]) failed_idx[w++] = s->first + i;
        }
    }
    return CAFE_B200_OK;
}

int cafe_b200_fetch_category_likelihoods(cafe_b200_ctx* c, int k, double* cat_lk)
{
    if (!c || !cat_lk || k < 1 || k > c->cap_k) return fail(c, CAFE_B200_ERR_ARG, "bad argument to fetch_category_likelihoods");
    for (Shard* s : c->shards) {
        if (s->n_families == 0) continue;
        CUDA_TRY(c, cudaSetDevice(s->device));
        CUDA_TRY(c, cudaMemcpyAsync(cat_lk + (size_t)s->first * k, s->d_cat_lk, (size_t)s->n_families * k * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    }
    return sync_all(c);
}

int cafe_b200_matrix_size(const cafe_b200_ctx* c) { return c ? c->n : 0; }

int cafe_b200_build_matrices(cafe_b200_ctx* c, const double* lambdas, int n_lambdas, int k, double* out)
{
    if (!c || !lambdas || !out || k < 1) return fail(c, CAFE_B200_ERR_ARG, "bad argument to build_matrices");
    std::vector<double> ones(c->n, 1.0);
    int rc = stage_and_build(c, lambdas, n_lambdas, k, nullptr, ones.data(), c->n);
    if (rc) return rc;
    const HostTree& t = c->tree;
    Shard* s = c->shards[0];
    CUDA_TRY(c, cudaSetDevice(s->device));
    std::vector<double> mt(c->mt_len);
    rc = sync_all(c);
    if (rc) return rc;
    const int* mat_of = reinterpret_cast<const int*>(c->h_param + c->lay.mat_of);
    const int cols = c->mf + 1;
    for (int cat = 0; cat < k; ++cat)
        for (int v = 0; v < t.n_nodes; ++v) {
            double* dst = out + ((size_t)cat * t.n_nodes + v) * c->n * cols;
            if (t.parent[v] < 0) { std::fill(dst, dst + (size_t)c->n * cols, 0.0); continue; }
            CUDA_TRY(c, cudaMemcpy(mt.data(), s->d_mt + (size_t)mat_of[cat * t.n_nodes + v] * c->mt_stride, c->mt_len * sizeof(double), cudaMemcpyDeviceToHost));
            for (int sz = 0; sz < c->n; ++sz)
                for (int cc = 0; cc < cols; ++cc) dst[(size_t)sz * cols + cc] = mt[(size_t)cc * c->nr + sz];
        }
    return CAFE_B200_OK;
}

int cafe_b200_prune_roots(cafe_b200_ctx* c, const double* lambdas, int n_lambdas, int k, double* out)
{
    if (!c || !lambdas || !out || k < 1) return fail(c, CAFE_B200_ERR_ARG, "bad argument to prune_roots");
    int rc = check_counts(c);
    if (rc) return rc;
    std::vector<double> ones(c->n, 1.0), cp(k, 1.0);
    rc = stage_and_build(c, lambdas, n_lambdas, k, cp.data(), ones.data(), c->n);
    if (rc) return rc;
    std::vector<double*> d_root(c->shards.size(), nullptr);
    for (size_t i = 0; i < c->shards.size() && rc == CAFE_B200_OK; ++i) {
        Shard* s = c->shards[i];
        const size_t total = (size_t)s->n_families * k * c->mrf;
        cudaError_t e = cudaSetDevice(s->device);
        if (e == cudaSuccess) e = dev_alloc(&d_root[i], total);
        if (e != cudaSuccess) { rc = fail(c, CAFE_B200_ERR_CUDA, cudaGetErrorString(e)); break; }
        rc = launch_prune(c, s, k, CAFE_B200_GAMMA_LINSUM, d_root[i]);
        if (rc == CAFE_B200_OK) {
            e = cudaEventRecord(s->ev[2], s->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(out + (size_t)s->first * k * c->mrf, d_root[i], total * sizeof(double), cudaMemcpyDeviceToHost, s->stream);
            if (e != cudaSuccess) rc = fail(c, CAFE_B200_ERR_CUDA, cudaGetErrorString(e));
            s->ev_valid[2] = true; s->ev_valid[3] = s->ev_valid[4] = false;
        }
    }
    const int rc2 = sync_all(c);
    for (size_t i = 0; i < c->shards.size(); ++i) {
        cudaSetDevice(c->shards[i]->device);
        cudaFree(d_root[i]);
    }
    return rc ? rc : rc2;
}

int cafe_b200_root_max(cafe_b200_ctx* c, const double* lambdas, int n_lambdas, double* out)
{
    if (!c || !lambdas || !out) return fail(c, CAFE_B200_ERR_ARG, "bad argument to root_max");
    NvtxRange r("cafe_b200_root_max");
    int rc = check_counts(c);
    if (rc) return rc;
    std::vector<double> ones(c->n, 1.0);
    const double cp = 1.0;
    rc = stage_and_build(c, lambdas, n_lambdas, 1, &cp, ones.data(), c->n);
    if (rc) return rc;
    for (Shard* s : c->shards) {
        CUDA_TRY(c, cudaSetDevice(s->device));
        rc = launch_prune(c, s, 1, CAFE_B200_ROOT_MAX, nullptr);
        if (rc) return rc;
        CUDA_TRY(c, cudaEventRecord(s->ev[2], s->stream));
        s->ev_valid[2] = true; s->ev_valid[3] = s->ev_valid[4] = false;
        if (s->n_families)
            CUDA_TRY(c, cudaMemcpyAsync(out + s->first, s->d_cat_km, (size_t)s->n_families * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    }
    rc = sync_all(c);
    if (rc) return rc;
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
    NvtxRange r("cafe_b200_branch_probabilities");
    const int nn = c->tree.n_nodes;
    const size_t total_all = (size_t)c->n_families * nn;
    const int hi = std::min(c->mf, c->n - 1);
    for (size_t i = 0; i < total_all; ++i)
        if (node_sizes[i] < 0 || node_sizes[i] > hi) return fail(c, CAFE_B200_ERR_COUNT_RANGE, "a node size is outside 0..max_family_size");
    std::vector<double> ones(c->n, 1.0);
    const double cp = 1.0;
    int rc = stage_and_build(c, lambdas, n_lambdas, 1, &cp, ones.data(), c->n);
    if (rc) return rc;
    const size_t ns = c->shards.size();
    std::vector<int32_t*> d_sizes(ns, nullptr);
    std::vector<uint8_t*> d_sel(ns, nullptr);
    std::vector<double*> d_out(ns, nullptr);
    cudaError_t e = cudaSuccess;
    for (size_t i = 0; i < ns && e == cudaSuccess; ++i) {
        Shard* s = c->shards[i];
        if (s->n_families == 0) continue;
        const size_t total = (size_t)s->n_families * nn;
        e = cudaSetDevice(s->device);
        if (e == cudaSuccess) e = dev_alloc(&d_sizes[i], total);
        if (e == cudaSuccess) e = dev_alloc(&d_out[i], total);
        if (e == cudaSuccess && selected) e = dev_alloc(&d_sel[i], (size_t)s->n_families);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_sizes[i], node_sizes + (size_t)s->first * nn, total * sizeof(int32_t), cudaMemcpyHostToDevice, s->stream);
        if (e == cudaSuccess && selected) e = cudaMemcpyAsync(d_sel[i], selected + s->first, (size_t)s->n_families, cudaMemcpyHostToDevice, s->stream);
        if (e == cudaSuccess) {
            const int blocks = (int)((total + VT_THREADS - 1) / VT_THREADS);
            viterbi_kernel<<<blocks, VT_THREADS, 0, s->stream>>>(s->n_families, nn, c->mf, c->nr, s->d_parent,
                                                                  reinterpret_cast<const int*>(s->d_param + c->lay.mat_of), s->d_mt, c->mt_stride,
                                                                  d_sizes[i], d_sel[i], d_out[i]);
            e = cudaGetLastError();
            c->launches++;
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(out + (size_t)s->first * nn, d_out[i], total * sizeof(double), cudaMemcpyDeviceToHost, s->stream);
    }
    for (size_t i = 0; i < ns; ++i) {
        Shard* s = c->shards[i];
        cudaSetDevice(s->device);
        cudaError_t e2 = cudaStreamSynchronize(s->stream);
        if (e == cudaSuccess) e = e2;
        cudaFree(d_sizes[i]); cudaFree(d_sel[i]); cudaFree(d_out[i]);
    }
    if (e != cudaSuccess) return fail(c, CAFE_B200_ERR_CUDA, std::string("branch_probabilities: ") + cudaGetErrorString(e));
    return CAFE_B200_OK;
}

int cafe_b200_reconstruct(cafe_b200_ctx* c, const double* lambdas, int n_lambdas, int k, const double* prior_by_size, int32_t* states)
{
    if (!c || !lambdas || !prior_by_size || !states || k < 1) return fail(c, CAFE_B200_ERR_ARG, "bad argument to reconstruct");
    if (c->hw_slots < 2) return fail(c, CAFE_B200_ERR_LIMIT, "reconstruction supports matrix sizes up to 256");
    NvtxRange r("cafe_b200_reconstruct");
    int rc = check_counts(c);
    if (rc) return rc;
    const int lim = std::min(c->mf, c->mrf) + 1;
    rc = stage_and_build(c, lambdas, n_lambdas, k, nullptr, prior_by_size, lim);
    if (rc) return rc;
    for (Shard* s : c->shards) {
        CUDA_TRY(c, cudaSetDevice(s->device));
        rc = launch_pupko(c, s, k);
        if (rc) return rc;
        if (s->n_families == 0) continue;
        CUDA_TRY(c, cudaEventRecord(s->ev[4], s->stream));
        const size_t n_states = (size_t)s->n_families * k * c->tree.n_internal;
        CUDA_TRY(c, cudaMemcpyAsync(states + (size_t)s->first * k * c->tree.n_internal, s->d_states, n_states * sizeof(int32_t), cudaMemcpyDeviceToHost, s->stream));
        s->ev_valid[2] = s->ev_valid[3] = false;
        s->ev_valid[4] = true;
    }
    return sync_all(c);
}

int64_t cafe_b200_launch_count(const cafe_b200_ctx* c) { return c ? c->launches : 0; }

int cafe_b200_last_timings(const cafe_b200_ctx* c, double* ms4)
{
    if (!c || !ms4) return CAFE_B200_ERR_ARG;
    for (int i = 0; i < 4; ++i) ms4[i] = 0.0;
    for (const Shard* s : c->shards) {
        float ms = 0.f;
        if (s->ev_valid[0] && s->ev_valid[1] && cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]) == cudaSuccess) ms4[0] = std::max(ms4[0], (double)ms);
        if (s->ev_valid[1] && s->ev_valid[2] && cudaEventElapsedTime(&ms, s->ev[1], s->ev[2]) == cudaSuccess) ms4[1] = std::max(ms4[1], (double)ms);
        if (s->ev_valid[2] && s->ev_valid[3] && cudaEventElapsedTime(&ms, s->ev[2], s->ev[3]) == cudaSuccess) ms4[2] = std::max(ms4[2], (double)ms);
        if (s->ev_valid[1] && s->ev_valid[4] && cudaEventElapsedTime(&ms, s->ev[1], s->ev[4]) == cudaSuccess) ms4[3] = std::max(ms4[3], (double)ms);
        cudaGetLastError();
    }
    return CAFE_B200_OK;
}

int cafe_b200_host_seconds(const cafe_b200_ctx* c, double* s3)
{
    if (!c || !s3) return CAFE_B200_ERR_ARG;
    for (int i = 0; i < 3; ++i) s3[i] = c->host_seconds[i];
    return CAFE_B200_OK;
}

int cafe_b200_timing_history(const cafe_b200_ctx* c, int n, double* ms)
{
    if (!c || !ms || n < 1) return CAFE_B200_ERR_ARG;
    n = std::min(n, TIMING_HISTORY);
    for (int j = 0; j < n; ++j) {
        double* row = ms + 4 * j;
        row[0] = row[1] = row[2] = row[3] = 0.0;
        for (const Shard* s : c->shards) {
            const int slot = ((s->hist_cur - j) % TIMING_HISTORY + TIMING_HISTORY) % TIMING_HISTORY;
            const cudaEvent_t* ev = s->hist_ev[slot];
            const bool* ok = s->hist_valid[slot];
            float t = 0.f;
            if (ok[0] && ok[1] && cudaEventElapsedTime(&t, ev[0], ev[1]) == cudaSuccess) row[0] = std::max(row[0], (double)t);
            if (ok[1] && ok[2] && cudaEventElapsedTime(&t, ev[1], ev[2]) == cudaSuccess) row[1] = std::max(row[1], (double)t);
            if (ok[2] && ok[3] && cudaEventElapsedTime(&t, ev[2], ev[3]) == cudaSuccess) row[2] = std::max(row[2], (double)t);
            if (ok[1] && ok[4] && cudaEventElapsedTime(&t, ev[1], ev[4]) == cudaSuccess) row[3] = std::max(row[3], (double)t);
            cudaGetLastError();
        }
    }
    return n;
}

int cafe_b200_describe(const cafe_b200_ctx* c, char* out, int cap)
{
    if (!c || !out || cap < 1) return CAFE_B200_ERR_ARG;
    snprintf(out, cap,
             "devices=%zu matrix=%d rows=%d geom(rb=%d,gw=%d,groups=%d,cps=%d,producers=%d) stages=%d smem=%d count_bytes=%d counts_in_smem=%d "
             "program_in_smem=%d peer_access=%d build_partitions=%d stack_depth=%d tmem_entries=%d tmem_cols=%d spill_entries=%d gemm_ops=%d ops=%zu",
             c->shards.size(), c->n, c->nr, c->geom.rb, c->geom.gw, c->geom.ng, c->geom.cps, c->geom.pw, c->n_stages, c->prune_smem, c->cnt_width,
             c->cnt_smem_bytes > 0 ? 1 : 0, c->ops_smem_bytes > 0 ? 1 : 0, c->peer_ok ? 1 : 0, c->ext_parts > 1 ? c->ext_parts : (c->peer_ok ? (int)c->shards.size() : 1),
             c->prog.depth, c->tmem_entries, c->tmem_cols, c->n_gspill, c->prog.n_gemm, c->prog.ops.size());
    return CAFE_B200_OK;
}

}  // extern "C"
