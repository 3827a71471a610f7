// Kernel 2+3 — fused Felsenstein-style pruning of a tile of FT families through the whole species
// tree, root prior / category weighting included.
//
// Restates (file:line in the reference)
//   inference_prune                          src/core.cpp:133-144
//   compute_node_probability                 src/probability.cpp:173-242   (leaf one-hot / error stencil,
//                                                                           internal = prod_children M_child * v_child,
//                                                                           root rows 1..mrf)
//   matrix::multiply                         src/matrix_cache.cpp:28-57
//   base_model root max of log L + log prior src/base_model.cpp:89-106
//   gamma_model::prune                       src/gamma_core.cpp:144-166
//
// Design (B200): one persistent thread block per SM walks a host-built post-order schedule for a
// tile of FT=32 families of one rate category.  Partial-likelihood vectors never leave the SM:
// they sit in shared-memory slots V[family][size].  An internal edge is the dense FP64 contraction
//      Y[NR x FT] = M_edge[NR x K] * V_child[K x FT]
// issued as mma.sync.m8n8k4.f64 (DMMA) by 8 consumer warps (4 row groups x 2 family groups, 40x16
// outputs per warp for NR=160), accumulators in registers for the whole K loop.  The matrix streams
// from L2 through a 4-stage shared-memory ring filled by 1-D bulk async copies (TMA engine,
// cp.async.bulk + mbarrier complete_tx) issued by a dedicated producer warp that runs ahead across
// ops, so the next edge's first chunks land while the consumers do leaf gathers or the epilogue.
// The epilogue multiplies the product straight into the parent's accumulator slot (child product);
// leaf edges are gathers of one matrix column (or an error-model stencil of columns), not GEMMs.
// HBM traffic per family is just its leaf counts in and k+1 doubles out.
#pragma once

#include "common.cuh"

namespace cafe {

template <int MB>
struct PruneSmem {
    static constexpr int NR = nr_of(MB);
    static constexpr int LDV = ldv_of(MB);
    static constexpr int STAGE_DOUBLES = stage_doubles(MB);
    static constexpr int STAGE_BYTES = STAGE_DOUBLES * 8;
    static constexpr int SLOT_DOUBLES = FT * LDV;
    static constexpr int SLOT_BYTES = SLOT_DOUBLES * 8;
    static constexpr int MISC_BYTES = 512;     // mbarriers
    static constexpr int EXP_BYTES = MAX_SLOTS * FT * 4;
    __host__ __device__ static constexpr int ring_bytes(int stages) { return stages * STAGE_BYTES; }
    __host__ __device__ static constexpr int total_bytes(int slots, int stages) { return ring_bytes(stages) + slots * SLOT_BYTES + CNT_CAP_BYTES + MISC_BYTES + EXP_BYTES; }
    __host__ static int max_slots(int smem_limit, int stages)
    {
        int s = (smem_limit - ring_bytes(stages) - CNT_CAP_BYTES - MISC_BYTES - EXP_BYTES) / SLOT_BYTES;
        return s > MAX_SLOTS ? MAX_SLOTS : s;
    }
};

template <int MB>
__global__ void __launch_bounds__(PRUNE_THREADS, 1) prune_kernel(const PruneParams p)
{
    using L = PruneSmem<MB>;
    constexpr int NR = L::NR;
    constexpr int LDV = L::LDV;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* ring = reinterpret_cast<double*>(smem_raw);
    const int ring_bytes = L::ring_bytes(p.n_stages);
    const uint32_t stage_mask = (uint32_t)p.n_stages - 1u;
    double* slots = reinterpret_cast<double*>(smem_raw + ring_bytes);
    uint16_t* cnt_s = reinterpret_cast<uint16_t*>(smem_raw + ring_bytes + p.n_slots * L::SLOT_BYTES);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + ring_bytes + p.n_slots * L::SLOT_BYTES + CNT_CAP_BYTES);
    uint64_t* empty_bar = full_bar + MAX_STAGES;
    int* slot_exp = reinterpret_cast<int*>(smem_raw + ring_bytes + p.n_slots * L::SLOT_BYTES + CNT_CAP_BYTES + L::MISC_BYTES);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < p.n_stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], CONSUMER_WARPS);
        }
        fence_barrier_init();
    }
    __syncthreads();

    const int64_t n_items = p.n_tiles * p.n_categories;

    if (warp == CONSUMER_WARPS) {
        // ===================== producer: stream matrix K-chunks into the ring =====================
        if (lane == 0) {
            uint32_t pos = 0;
            for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int cat = (int)(item / p.n_tiles);
                const POp* ops = p.ops + (size_t)cat * p.n_ops;
                for (int o = 0; o < p.n_ops; ++o) {
                    const int type = ops[o].type;
                    if (type != OP_GEMM_SET && type != OP_GEMM_MUL && type != OP_GEMM_SET_LEAF && type != OP_GEMM_MUL_LEAF) continue;
                    const double* src = p.mp + (size_t)ops[o].mat * p.mp_stride;
                    for (int ch = 0; ch < p.n_kchunks; ++ch, ++pos) {
                        const uint32_t stage = pos & stage_mask;
                        const uint32_t round = pos >> p.stage_shift;
                        mbar_wait(&empty_bar[stage], (round & 1) ^ 1);
                        mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
                        bulk_copy_g2s(ring + (size_t)stage * L::STAGE_DOUBLES, src + (size_t)ch * L::STAGE_DOUBLES, L::STAGE_BYTES, &full_bar[stage]);
                    }
                }
            }
        }
        return;
    }

    // =============================== consumers =====================================================
    const int warp_m = warp & 3;          // row group: rows [warp_m*8*MB, +8*MB)
    const int warp_n = warp >> 2;         // family group: columns [warp_n*16, +16)
    const int g = lane >> 2;              // fragment row / column group
    const int t4 = lane & 3;
    uint32_t pos = 0;

    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int cat = (int)(item / p.n_tiles);
        const int64_t tile = item % p.n_tiles;
        const int64_t fam0 = tile * FT;

        consumer_sync();      // previous item fully finished with shared memory
        if (p.counts_in_smem) {
            const int total = FT * p.n_leaves;
            for (int i = tid; i < total; i += CONSUMER_THREADS) {
                const int f = i / p.n_leaves;
                int64_t fam = fam0 + f;
                if (fam >= p.n_families) fam = p.n_families - 1;
                cnt_s[i] = (uint16_t)p.counts[fam * p.n_leaves + (i - f * p.n_leaves)];
            }
        }
        if (tid < MAX_SLOTS * FT) slot_exp[tid] = 0;
        consumer_sync();

        const POp* ops = p.ops + (size_t)cat * p.n_ops;
        POp next = ops[0];
        for (int o = 0; o < p.n_ops; ++o) {
            const POp op = next;
            if (o + 1 < p.n_ops) next = ops[o + 1];          // prefetch: the load overlaps this op's work
            switch (op.type) {
            case OP_LEAF_SET2: {
                // ---- a cherry in one pass: V = column(leaf 1) * column(leaf 2) ----
                const double* mt1 = p.mt + (size_t)op.mat * p.mt_stride;
                const double* mt2 = p.mt + (size_t)op.mat2 * p.mt_stride;
                double* dst = slots + (size_t)op.a * L::SLOT_DOUBLES;
                double v1[FT / CONSUMER_WARPS][MB], v2[FT / CONSUMER_WARPS][MB];
                #pragma unroll
                for (int fi = 0; fi < FT / CONSUMER_WARPS; ++fi) {
                    const int f = warp * (FT / CONSUMER_WARPS) + fi;
                    int o1, o2;
                    if (p.counts_in_smem) { o1 = cnt_s[f * p.n_leaves + op.col]; o2 = cnt_s[f * p.n_leaves + op.col2]; }
                    else {
                        int64_t fam = fam0 + f;
                        if (fam >= p.n_families) fam = p.n_families - 1;
                        o1 = p.counts[fam * p.n_leaves + op.col]; o2 = p.counts[fam * p.n_leaves + op.col2];
                    }
                    #pragma unroll
                    for (int i = 0; i < MB; ++i) {
                        v1[fi][i] = __ldg(mt1 + (size_t)o1 * NR + lane + 32 * i);
                        v2[fi][i] = __ldg(mt2 + (size_t)o2 * NR + lane + 32 * i);
                    }
                }
                #pragma unroll
                for (int fi = 0; fi < FT / CONSUMER_WARPS; ++fi) {
                    const int f = warp * (FT / CONSUMER_WARPS) + fi;
                    double* row = dst + (size_t)f * LDV;
                    #pragma unroll
                    for (int i = 0; i < MB; ++i) row[lane + 32 * i] = v1[fi][i] * v2[fi][i];
                    if (lane == 0) slot_exp[op.a * FT + f] = 0;
                }
                consumer_sync();
                break;
            }
            case OP_LEAF_SET:
            case OP_LEAF_MUL: {
                // ---- leaf edge: gather column obs (or an error-model stencil of columns) of M^T ----
                const double* mt = p.mt + (size_t)op.mat * p.mt_stride;
                const int col = op.col;
                double* dst = slots + (size_t)op.a * L::SLOT_DOUBLES;
                #pragma unroll
                for (int fi = 0; fi < FT / CONSUMER_WARPS; ++fi) {
                    const int f = warp * (FT / CONSUMER_WARPS) + fi;
                    int obs;
                    if (p.counts_in_smem) obs = cnt_s[f * p.n_leaves + col];
                    else {
                        int64_t fam = fam0 + f;
                        if (fam >= p.n_families) fam = p.n_families - 1;
                        obs = p.counts[fam * p.n_leaves + col];
                    }
                    double v[MB];
                    if (p.err == nullptr) {
                        const double* src = mt + (size_t)obs * NR;
                        #pragma unroll
                        for (int i = 0; i < MB; ++i) v[i] = __ldg(src + lane + 32 * i);
                    }
                    else {
                        // y[s] = sum over deviations, ascending child size, no FMA (src/probability.cpp:182-193
                        // feeding src/matrix_cache.cpp:48-54)
                        #pragma unroll
                        for (int i = 0; i < MB; ++i) v[i] = 0.0;
                        const int offset = obs - (p.err_ndev - 1) / 2;
                        for (int d = 0; d < p.err_ndev; ++d) {
                            const int c = offset + d;
                            if (c < 0 || c > p.mf) continue;
                            const double pe = __ldg(p.err + (size_t)obs * p.err_ndev + d);
                            const double* src = mt + (size_t)c * NR;
                            #pragma unroll
                            for (int i = 0; i < MB; ++i) v[i] = __dadd_rn(v[i], __dmul_rn(__ldg(src + lane + 32 * i), pe));
                        }
                    }
                    double* row = dst + (size_t)f * LDV;
                    if (op.type == OP_LEAF_SET) {
                        #pragma unroll
                        for (int i = 0; i < MB; ++i) row[lane + 32 * i] = v[i];
                        if (lane == 0) slot_exp[op.a * FT + f] = 0;     // fresh vector in a recycled slot
                    }
                    else {
                        #pragma unroll
                        for (int i = 0; i < MB; ++i) row[lane + 32 * i] *= v[i];
                    }
                }
                consumer_sync();
                break;
            }
            case OP_GEMM_SET:
            case OP_GEMM_MUL:
            case OP_GEMM_SET_LEAF:
            case OP_GEMM_MUL_LEAF: {
                // ---- internal edge: Y = M * V_child on the FP64 tensor pipe ----
                const bool is_set = (op.type == OP_GEMM_SET || op.type == OP_GEMM_SET_LEAF);
                const bool with_leaf = (op.type == OP_GEMM_SET_LEAF || op.type == OP_GEMM_MUL_LEAF);
                const int src_slot = is_set ? op.a : op.b;
                const double* vsrc = slots + (size_t)src_slot * L::SLOT_DOUBLES + (size_t)(warp_n * 16 + g) * LDV + t4;
                double acc[MB][2][2];
                double lf[MB][2][2];       // leaf-sibling factor of each output element (1.0 when there is none)
                #pragma unroll
                for (int i = 0; i < MB; ++i) {
                    acc[i][0][0] = acc[i][0][1] = acc[i][1][0] = acc[i][1][1] = 0.0;
                    lf[i][0][0] = lf[i][0][1] = lf[i][1][0] = lf[i][1][1] = 1.0;
                }
                if (with_leaf) {
                    // issued before the K loop so the L2 latency of the gather hides behind the MMAs
                    const double* mt2 = p.mt + (size_t)op.mat2 * p.mt_stride + warp_m * 8 * MB + g;
                    #pragma unroll
                    for (int nb = 0; nb < 2; ++nb)
                        #pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int f = warp_n * 16 + nb * 8 + t4 * 2 + e;
                            int obs;
                            if (p.counts_in_smem) obs = cnt_s[f * p.n_leaves + op.col2];
                            else {
                                int64_t fam = fam0 + f;
                                if (fam >= p.n_families) fam = p.n_families - 1;
                                obs = p.counts[fam * p.n_leaves + op.col2];
                            }
                            #pragma unroll
                            for (int i = 0; i < MB; ++i) lf[i][nb][e] = __ldg(mt2 + (size_t)obs * NR + i * 8);
                        }
                }
                const int a_off = (warp_m * 8 * MB) * 4 + lane;
                for (int ch = 0; ch < p.n_kchunks; ++ch, ++pos) {
                    const uint32_t stage = pos & stage_mask;
                    mbar_wait(&full_bar[stage], (pos >> p.stage_shift) & 1);
                    const double* a_stage = ring + (size_t)stage * L::STAGE_DOUBLES + a_off;
                    #pragma unroll
                    for (int pp = 0; pp < PPS; ++pp) {
                        const int kcol = (ch * PPS + pp) * 4;
                        double a[MB];
                        #pragma unroll
                        for (int i = 0; i < MB; ++i) a[i] = a_stage[pp * NR * 4 + i * 32];
                        const double b0 = vsrc[kcol];
                        const double b1 = vsrc[8 * LDV + kcol];
                        #pragma unroll
                        for (int i = 0; i < MB; ++i) {
                            dmma_m8n8k4(acc[i][0][0], acc[i][0][1], a[i], b0);
                            dmma_m8n8k4(acc[i][1][0], acc[i][1][1], a[i], b1);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_bar[stage]);
                }
                if (is_set) consumer_sync();      // in place: every warp is done reading V_child before anyone overwrites it
                double* dst = slots + (size_t)op.a * L::SLOT_DOUBLES;
                #pragma unroll
                for (int i = 0; i < MB; ++i) {
                    const int s = warp_m * 8 * MB + i * 8 + g;
                    #pragma unroll
                    for (int nb = 0; nb < 2; ++nb) {
                        const int f = warp_n * 16 + nb * 8 + t4 * 2;
                        double* q0 = dst + (size_t)f * LDV + s;
                        double* q1 = q0 + LDV;
                        const double y0 = acc[i][nb][0] * lf[i][nb][0];
                        const double y1 = acc[i][nb][1] * lf[i][nb][1];
                        if (is_set) { *q0 = y0; *q1 = y1; }
                        else { *q0 *= y0; *q1 *= y1; }
                    }
                }
                if (!is_set && p.rescale && tid < FT) slot_exp[op.a * FT + tid] += slot_exp[op.b * FT + tid];
                consumer_sync();
                break;
            }
            case OP_SPILL:
            case OP_FILL: {
                double* sl = slots + (size_t)op.a * L::SLOT_DOUBLES;
                double* sc = p.scratch + ((size_t)blockIdx.x * p.n_spill + op.b) * L::SLOT_DOUBLES;
                int* sce = p.scratch_exp + ((size_t)blockIdx.x * p.n_spill + op.b) * FT;
                if (op.type == OP_SPILL) {
                    for (int i = tid; i < L::SLOT_DOUBLES; i += CONSUMER_THREADS) sc[i] = sl[i];
                    if (tid < FT) sce[tid] = slot_exp[op.a * FT + tid];
                }
                else {
                    for (int i = tid; i < L::SLOT_DOUBLES; i += CONSUMER_THREADS) sl[i] = sc[i];
                    if (tid < FT) slot_exp[op.a * FT + tid] = sce[tid];
                }
                consumer_sync();
                break;
            }
            case OP_RESCALE: {
                // exact power-of-two renormalisation per family: V *= 2^-e, exponent tracked in slot_exp
                if (!p.rescale) break;          // uniform: reference arithmetic, no renormalisation
                double* sl = slots + (size_t)op.a * L::SLOT_DOUBLES;
                #pragma unroll
                for (int fi = 0; fi < FT / CONSUMER_WARPS; ++fi) {
                    const int f = warp * (FT / CONSUMER_WARPS) + fi;
                    double* row = sl + (size_t)f * LDV;
                    double m = 0.0;
                    for (int s = lane; s <= p.mf; s += 32) m = fmax(m, row[s]);
                    #pragma unroll
                    for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
                    if (m > 0.0 && m < 0x1p-64) {
                        int e;
                        frexp(m, &e);                       // m = f * 2^e, f in [0.5, 1)
                        const double sc = ldexp(1.0, -e);   // exact; brings the max into [0.5, 1)
                        for (int s = lane; s < NR; s += 32) row[s] *= sc;
                        if (lane == 0) slot_exp[op.a * FT + f] += e;
                    }
                }
                consumer_sync();
                break;
            }
            case OP_ROOT: {
                // ---- root: index j <-> root size j+1 (src/base_model.cpp:95-98) ----
                const double* sl = slots + (size_t)op.a * L::SLOT_DOUBLES;
                #pragma unroll
                for (int fi = 0; fi < FT / CONSUMER_WARPS; ++fi) {
                    const int f = warp * (FT / CONSUMER_WARPS) + fi;
                    const int64_t fam = fam0 + f;
                    if (fam >= p.n_families) continue;      // warp-uniform
                    const double* row = sl + (size_t)f * LDV;
                    const int e = p.rescale ? slot_exp[op.a * FT + f] : 0;
                    if (p.root_out) {
                        double* out = p.root_out + ((size_t)fam * p.n_categories + cat) * p.mrf;
                        for (int j = lane; j < p.mrf; j += 32) out[j] = e ? ldexp(row[j + 1], e) : row[j + 1];
                    }
                    if (p.mode == 0) {
                        // max_j( log L[j] + log prior[j] )
                        double best = -INFINITY;
                        for (int j = lane; j < p.mrf; j += 32) {
                            double lj = log(row[j + 1]);
                            if (e) lj += (double)e * 0.69314718055994530942;
                            best = fmax(best, lj + p.logprior[j]);
                        }
                        #pragma unroll
                        for (int off = 16; off > 0; off >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, off));
                        if (lane == 0) { p.cat_lk[fam] = best; p.fail[fam] = 0; }
                    }
                    else {
                        // fail when every root entry is zero; else max_j( L[j] * prior[j] ) * catprob
                        double best = 0.0;
                        int nonzero = 0;
                        for (int j = lane; j < p.mrf; j += 32) {
                            double lj = row[j + 1];
                            if (e) lj = ldexp(lj, e);
                            nonzero |= (lj != 0.0);
                            const double full = lj * p.prior[j];
                            if (best < full) best = full;
                        }
                        #pragma unroll
                        for (int off = 16; off > 0; off >>= 1) {
                            best = fmax(best, __shfl_xor_sync(0xffffffffu, best, off));
                            nonzero |= __shfl_xor_sync(0xffffffffu, nonzero, off);
                        }
                        if (lane == 0) {
                            p.cat_lk[fam * p.n_categories + cat] = best * p.cat_probs[cat];
                            p.fail[fam * p.n_categories + cat] = nonzero ? 0 : 1;
                        }
                    }
                }
                break;
            }
            default:
                break;
            }
        }
    }
}

}  // namespace cafe
