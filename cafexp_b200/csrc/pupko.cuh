// Kernel 4 — Pupko joint ancestral reconstruction (max-product with argmax tables and traceback).
//
// Restates (file:line in the reference)
//   reconstruct_leaf_node       src/gene_family_reconstructor.cpp:13-33   L[i] = M(i, obs), i >= 1; L[0] = 0; no error model
//   reconstruct_internal_node   src/gene_family_reconstructor.cpp:74-112  L[i] = max_j (prod_children L_c[j]) * M(i, j), strict '>' from -1
//   reconstruct_root_node       src/gene_family_reconstructor.cpp:35-72   argmax_{j>=1} (prod_children L_c[j]) * prior(j)
//   reconstruct_gene_family     src/gene_family_reconstructor.cpp:131-165 traceback state[child] = C_child[state[parent]]
//
// Same skeleton as the pruning kernel (persistent blocks, a tile of FT families per block, vectors
// in shared-memory slots, matrix K-chunks streamed through a bulk-copy ring by a producer warp, the
// same host schedule), but the contraction is over the (max, x) semiring with an argmax, which the
// tensor cores cannot do: it runs on the FP64 ALU with 4 families x MB rows per thread in registers,
// the child index j strictly ascending inside each thread so that "first maximum wins" holds exactly.
// Products are formed in the reference's order (children in Newick order, then the matrix entry)
// without FMA, so the integer result can only differ from the reference through the <1 ulp
// difference between CUDA's and glibc's exp() in the matrix entries.
// Bound: FP64 ALU issue (DMUL + DSETP per element); reported as elements/s, not against tensor peak.
#pragma once

#include "common.cuh"

namespace cafe {

struct PupkoParams {
    int64_t n_families;
    int n_leaves;
    int n_nodes;
    int n_internal;
    int n_categories;
    int mf;
    int mrf;
    int n_ops;
    int n_kchunks;
    int n_spill;
    int n_slots;
    int counts_in_smem;
    int cnt_width;                  // bytes per leaf count: 1 or 2
    int64_t n_tiles;
    const Op* ops;
    const void* counts;
    const int* leaf_col;
    const int* parent;
    const int* internal_idx;        // [n_nodes] position among internal nodes, -1 for leaves
    const int* mat_of;
    const double* mt;
    size_t mt_stride;
    const double* prior;            // indexed by root size
    double* scratch;                // [grid][n_spill][FT*LDV]
    uint8_t* ctab;                  // [grid][n_internal][FT][NR] argmax tables of the current tile (indexed by internal node:
                                    // half the footprint of a per-node table, so 148 tiles in flight stay L2-resident)
    int32_t* states;                // [F][k][n_internal]
};

template <int MB>
struct PupkoSmem {
    static constexpr int NR = nr_of(MB);
    static constexpr int LDV = ldv_of(MB);
    static constexpr int STAGE_DOUBLES = stage_doubles(MB);     // PPS*4 = 8 matrix columns of NR rows
    static constexpr int STAGE_BYTES = STAGE_DOUBLES * 8;
    static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
    static constexpr int SLOT_DOUBLES = FT * LDV;
    static constexpr int SLOT_BYTES = SLOT_DOUBLES * 8;
    static constexpr int MISC_BYTES = 512;
    __host__ __device__ static constexpr int total_bytes(int slots) { return RING_BYTES + slots * SLOT_BYTES + CNT_CAP_BYTES + MISC_BYTES; }
};

constexpr int JC = PPS * 4;     // child sizes per ring stage
constexpr int PUPKO_WARPS = 8;           // consumer warps
__host__ __device__ constexpr int pupko_threads(int cw) { return (cw + 1) * 32; }     // + 1 producer warp

// One child size j against four families of one parent-size row: val_f = v_f * m, strict '>' keeps the first maximum
// (src/gene_family_reconstructor.cpp:96-105).  The four chains are written out with their own predicates so that the
// multiply -> compare -> move latencies of different families overlap (left to the compiler every pair went through
// one predicate register, i.e. one FP64 round trip per element: 35 % FP64 pipe use in round 1), and the update is a
// predicated move rather than a select: selects all issue on the half-rate ALU pipe (three per element, the bound once the
// latencies overlap), predicated moves can also go to the FMA pipe.
__device__ __forceinline__ void maxprod4(double& b0, double& b1, double& b2, double& b3, int& a0, int& a1, int& a2, int& a3, double v0, double v1,
                                         double v2, double v3, double m, int j)
{
    asm("{\n"
        ".reg .pred p0, p1, p2, p3;\n"
        ".reg .f64 t0, t1, t2, t3;\n"
        "mul.rn.f64 t0, %8, %12;\n"
        "mul.rn.f64 t1, %9, %12;\n"
        "mul.rn.f64 t2, %10, %12;\n"
        "mul.rn.f64 t3, %11, %12;\n"
        "setp.gt.f64 p0, t0, %0;\n"
        "setp.gt.f64 p1, t1, %1;\n"
        "setp.gt.f64 p2, t2, %2;\n"
        "setp.gt.f64 p3, t3, %3;\n"
        "@p0 mov.f64 %0, t0;\n"
        "@p1 mov.f64 %1, t1;\n"
        "@p2 mov.f64 %2, t2;\n"
        "@p3 mov.f64 %3, t3;\n"
        "@p0 mov.s32 %4, %13;\n"
        "@p1 mov.s32 %5, %13;\n"
        "@p2 mov.s32 %6, %13;\n"
        "@p3 mov.s32 %7, %13;\n"
        "}\n"
        : "+d"(b0), "+d"(b1), "+d"(b2), "+d"(b3), "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3)
        : "d"(v0), "d"(v1), "d"(v2), "d"(v3), "d"(m), "r"(j));
}

// Eight consumer warps (four families each) + one producer warp.  A 16-warp variant (two families per warp, 92 registers,
// four warps per sub-partition) measured the same 324.0 ms: the bound is not latency but the two half-rate pipes — FP64
// (multiply, compare) and ALU (two 64-bit selects + one 32-bit select per element) are together busy all the time
// (ncu: 42 % + 64 %), 10 cycles per element and sub-partition; ptxas turns predicated moves and multiply-add "moves"
// back into selects, so the selects cannot be moved to the FMA pipe.
template <int MB>
__global__ void __launch_bounds__(pupko_threads(PUPKO_WARPS), 1) pupko_kernel(const PupkoParams p)
{
    constexpr int CONSUMER_WARPS = PUPKO_WARPS;
    constexpr int CONSUMER_THREADS = PUPKO_WARPS * 32;
    auto consumer_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(PUPKO_WARPS * 32) : "memory"); };
    using L = PupkoSmem<MB>;
    constexpr int NR = L::NR;
    constexpr int LDV = L::LDV;
    constexpr int FPW = FT / CONSUMER_WARPS;      // families per warp
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* ring = reinterpret_cast<double*>(smem_raw);
    double* slots = reinterpret_cast<double*>(smem_raw + L::RING_BYTES);
    uint16_t* cnt_s = reinterpret_cast<uint16_t*>(smem_raw + L::RING_BYTES + p.n_slots * L::SLOT_BYTES);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + L::RING_BYTES + p.n_slots * L::SLOT_BYTES + CNT_CAP_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], CONSUMER_WARPS);
        }
        fence_barrier_init();
    }
    __syncthreads();

    const int64_t n_items = p.n_tiles * p.n_categories;

    if (warp == CONSUMER_WARPS) {
        if (lane == 0) {
            uint32_t pos = 0;
            for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int cat = (int)(item / p.n_tiles);
                for (int o = 0; o < p.n_ops; ++o) {
                    const Op op = p.ops[o];
                    if (op.type != OP_GEMM_SET && op.type != OP_GEMM_MUL) continue;
                    const double* src = p.mt + (size_t)p.mat_of[cat * p.n_nodes + op.node] * p.mt_stride;
                    for (int ch = 0; ch < p.n_kchunks; ++ch, ++pos) {
                        const uint32_t stage = pos % STAGES;
                        mbar_wait(&empty_bar[stage], ((pos / STAGES) & 1) ^ 1);
                        mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
                        bulk_copy_g2s(ring + (size_t)stage * L::STAGE_DOUBLES, src + (size_t)ch * L::STAGE_DOUBLES, L::STAGE_BYTES, &full_bar[stage]);
                    }
                }
            }
        }
        return;
    }

    uint32_t pos = 0;
    uint8_t* ctab = p.ctab + (size_t)blockIdx.x * p.n_internal * FT * NR;
    auto count_at = [&](int64_t i) -> int {
        return p.cnt_width == 1 ? (int)reinterpret_cast<const uint8_t*>(p.counts)[i] : (int)reinterpret_cast<const uint16_t*>(p.counts)[i];
    };

    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int cat = (int)(item / p.n_tiles);
        const int64_t tile = item % p.n_tiles;
        const int64_t fam0 = tile * FT;

        consumer_sync();
        if (p.counts_in_smem) {
            const int total = FT * p.n_leaves;
            for (int i = tid; i < total; i += CONSUMER_THREADS) {
                const int f = i / p.n_leaves;
                int64_t fam = fam0 + f;
                if (fam >= p.n_families) fam = p.n_families - 1;
                cnt_s[i] = (uint16_t)count_at(fam * p.n_leaves + (i - f * p.n_leaves));
            }
        }
        consumer_sync();

        for (int o = 0; o < p.n_ops; ++o) {
            const Op op = p.ops[o];
            switch (op.type) {
            case OP_LEAF_SET:
            case OP_LEAF_MUL: {
                const double* mt = p.mt + (size_t)p.mat_of[cat * p.n_nodes + op.node] * p.mt_stride;
                const int col = p.leaf_col[op.node];
                double* dst = slots + (size_t)op.a * L::SLOT_DOUBLES;
                #pragma unroll
                for (int fi = 0; fi < FPW; ++fi) {
                    const int f = warp * FPW + fi;
                    int obs;
                    if (p.counts_in_smem) obs = cnt_s[f * p.n_leaves + col];
                    else {
                        int64_t fam = fam0 + f;
                        if (fam >= p.n_families) fam = p.n_families - 1;
                        obs = count_at(fam * p.n_leaves + col);
                    }
                    const double* src = mt + (size_t)obs * NR;
                    double* row = dst + (size_t)f * LDV;
                    #pragma unroll
                    for (int i = 0; i < MB; ++i) {
                        const int s = lane + 32 * i;
                        double v = __ldg(src + s);
                        if (s == 0) v = 0.0;                          // L[0] is never assigned: stays 0
                        if (op.type == OP_LEAF_SET) row[s] = v;
                        else row[s] = __dmul_rn(row[s], v);
                    }
                }
                consumer_sync();
                break;
            }
            case OP_GEMM_SET:
            case OP_GEMM_MUL: {
                const int src_slot = (op.type == OP_GEMM_SET) ? op.a : op.b;
                const double* vsrc = slots + (size_t)src_slot * L::SLOT_DOUBLES + (size_t)(warp * FPW) * LDV;
                double best[FPW][MB];
                int arg[FPW][MB];
                #pragma unroll
                for (int fi = 0; fi < FPW; ++fi)
                    #pragma unroll
                    for (int i = 0; i < MB; ++i) { best[fi][i] = -1.0; arg[fi][i] = 0; }
                for (int ch = 0; ch < p.n_kchunks; ++ch, ++pos) {
                    const uint32_t stage = pos % STAGES;
                    mbar_wait(&full_bar[stage], (pos / STAGES) & 1);
                    const double* m_stage = ring + (size_t)stage * L::STAGE_DOUBLES + lane;
                    // Columns beyond max_family_size are zero in the transposed matrix (padding up to whole stages), so
                    // their products are 0 (or NaN against a stale slot entry): neither beats a maximum that is >= 0
                    // after j = 0 — no bounds test in the loop.
                    static_assert(FPW == 4, "maxprod4 handles the four families of a warp");
                    #pragma unroll
                    for (int jj = 0; jj < JC; ++jj) {
                        const int j = ch * JC + jj;
                        double m[MB];
                        #pragma unroll
                        for (int i = 0; i < MB; ++i) m[i] = m_stage[jj * NR + 32 * i];
                        const double v0 = vsrc[0 * LDV + j], v1 = vsrc[1 * LDV + j], v2 = vsrc[2 * LDV + j], v3 = vsrc[3 * LDV + j];
                        #pragma unroll
                        for (int i = 0; i < MB; ++i)
                            maxprod4(best[0][i], best[1][i], best[2][i], best[3][i], arg[0][i], arg[1][i], arg[2][i], arg[3][i], v0, v1, v2, v3, m[i], j);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_bar[stage]);
                }
                consumer_sync();
                double* dst = slots + (size_t)op.a * L::SLOT_DOUBLES;
                uint8_t* ct = ctab + (size_t)p.internal_idx[op.node] * FT * NR;
                #pragma unroll
                for (int fi = 0; fi < FPW; ++fi) {
                    const int f = warp * FPW + fi;
                    #pragma unroll
                    for (int i = 0; i < MB; ++i) {
                        const int s = lane + 32 * i;
                        double* q = dst + (size_t)f * LDV + s;
                        if (op.type == OP_GEMM_SET) *q = best[fi][i];
                        else *q = __dmul_rn(*q, best[fi][i]);
                        ct[(size_t)f * NR + s] = (uint8_t)arg[fi][i];
                    }
                }
                consumer_sync();
                break;
            }
            case OP_SPILL:
            case OP_FILL: {
                double* sl = slots + (size_t)op.a * L::SLOT_DOUBLES;
                double* sc = p.scratch + ((size_t)blockIdx.x * p.n_spill + op.b) * L::SLOT_DOUBLES;
                if (op.type == OP_SPILL) for (int i = tid; i < L::SLOT_DOUBLES; i += CONSUMER_THREADS) sc[i] = sl[i];
                else for (int i = tid; i < L::SLOT_DOUBLES; i += CONSUMER_THREADS) sl[i] = sc[i];
                consumer_sync();
                break;
            }
            case OP_ROOT: {
                // root state, then traceback; one thread per family (tables of this tile were written by this block)
                __threadfence_block();
                consumer_sync();
                if (tid < FT && fam0 + tid < p.n_families) {
                    const int f = tid;
                    const double* row = slots + (size_t)op.a * L::SLOT_DOUBLES + (size_t)f * LDV;
                    const int lim = (p.mf < p.mrf ? p.mf : p.mrf) + 1;
                    double max_val = -1.0;
                    int root_state = 0;
                    for (int j = 1; j < lim; ++j) {
                        const double val = __dmul_rn(row[j], p.prior[j]);
                        if (val > max_val) { max_val = val; root_state = j; }
                    }
                    int32_t* out = p.states + ((size_t)(fam0 + f) * p.n_categories + cat) * p.n_internal;
                    const int root = p.n_nodes - 1;
                    out[p.internal_idx[root]] = root_state;
                    for (int v = root - 1; v >= 0; --v) {
                        const int ii = p.internal_idx[v];
                        if (ii < 0) continue;
                        const int ps = out[p.internal_idx[p.parent[v]]];
                        out[ii] = ctab[((size_t)ii * FT + f) * NR + ps];
                    }
                }
                break;
            }
            default:
                break;      // OP_RESCALE: the reference never rescales here
            }
        }
    }
}

}  // namespace cafe
