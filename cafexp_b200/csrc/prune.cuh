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
// Design (B200): one persistent thread block per SM walks a host-built post-order schedule for a tile of
// NG*16 families of one rate category.  Partial-likelihood vectors never leave the SM: they sit in shared-memory
// slots V[family][size].  An internal edge is the dense FP64 contraction
//      Y[NR x 16] = M_edge[NR x K] * V_child[K x 16]
// issued as mma.sync.m8n8k4.f64 (DMMA), accumulators in registers for the whole K loop.  The matrix streams from
// L2 through a shared-memory ring filled by 1-D bulk async copies (TMA engine, cp.async.bulk + mbarrier
// complete_tx) issued by two producer warps (alternate chunks) that run ahead across ops and items.
//
// The consumer warps form NG = 2 GROUPS of 4 warps (one warp per SM sub-partition each).  Group g owns families
// [16g, 16g+16) of the tile (its rows of every slot), has its own named barrier and walks the same op list at its
// own pace; both groups consume the SAME matrix stream from the shared ring (a stage is released when every consumer
// warp has read it), so the L2 traffic per flop is that of a 32-family tile and the groups can drift up to one ring
// depth (4 stages of 2 chunks) apart: while one group is between GEMMs (epilogue, child product, leaf gathers at L2 latency,
// barriers) the other keeps the FP64 pipe busy.  Inside a GEMM the A/B fragments are double-buffered in registers
// across ring stages; the leaf-sibling factors of the epilogue are gathered after the K loop, when the fragment
// registers are free.  NG = 3 (48-family tiles, two slots) is compiled for experiments, see plan_shared_memory.
//
// What bounds the kernel (scripts/kloop_mix.cu rebuilds the loop from its parts on the same geometry): the bare
// loop runs at the DMMA peak; the per-stage mbarrier wait / arrive costs 11 %, the FP64 multiplies of the epilogue
// 4 %, a leaf gather at L2 latency per GEMM 11 % (the tree has one per three GEMMs), and op dispatch, count lookups,
// spills and the root reduction the rest: 0.71 of the DMMA peak in all (profiles/r01_pruning_kernel_experiments.md).
//
// The epilogue multiplies the product straight into the parent's accumulator slot (child product), fused with a
// leaf sibling's gathered column; leaf edges are gathers of one matrix column (or an error-model stencil of
// columns), not GEMMs.  HBM traffic per family is just its leaf counts in and k+1 doubles out.
#pragma once

#include "common.cuh"

namespace cafe {

constexpr int GROUP_WARPS = 4;                           // one per SM sub-partition
constexpr int GROUP_THREADS = GROUP_WARPS * 32;
constexpr int GFT = 16;                                  // families per group = 2 n8 blocks
constexpr int FPW = GFT / GROUP_WARPS;                   // 4 families per warp in the gather / root ops
constexpr int PRODUCER_WARPS = 2;                        // take alternate chunks of the one matrix stream
constexpr int MAX_GROUPS = 3;
constexpr int CPS = 2;                                   // K chunks (of PPS panels) per ring stage: one mbarrier pair per 40 DMMAs
constexpr int PRUNE_CNT_CAP_BYTES = 12288;               // staged leaf counts (uint16) of a 48-family tile, if they fit
__host__ __device__ constexpr int prune_threads(int ng) { return (ng * GROUP_WARPS + PRODUCER_WARPS) * 32; }

template <int MB, int NG>
struct PruneSmem {
    static constexpr int PFT = NG * GFT;                 // families per thread-block tile
    static constexpr int NR = nr_of(MB);
    static constexpr int LDV = ldv_of(MB);
    static constexpr int CHUNK_DOUBLES = stage_doubles(MB);          // PPS panels = 8 matrix columns
    static constexpr int CHUNK_BYTES = CHUNK_DOUBLES * 8;
    static constexpr int STAGE_DOUBLES = CPS * CHUNK_DOUBLES;        // a ring stage = CPS chunks (contiguous in the panelised matrix)
    static constexpr int STAGE_BYTES = STAGE_DOUBLES * 8;
    static constexpr int SLOT_DOUBLES = PFT * LDV;
    static constexpr int SLOT_BYTES = SLOT_DOUBLES * 8;
    static constexpr int MISC_BYTES = 512;     // mbarriers
    static constexpr int EXP_BYTES = MAX_SLOTS * PFT * 4;
    __host__ __device__ static constexpr int ring_bytes(int chunks) { return chunks * CHUNK_BYTES; }        // ring memory is sized in chunks
    __host__ __device__ static constexpr int total_bytes(int slots, int chunks) { return ring_bytes(chunks) + slots * SLOT_BYTES + PRUNE_CNT_CAP_BYTES + MISC_BYTES + EXP_BYTES; }
    __host__ static int max_slots(int smem_limit, int chunks)
    {
        int s = (smem_limit - ring_bytes(chunks) - PRUNE_CNT_CAP_BYTES - MISC_BYTES - EXP_BYTES) / SLOT_BYTES;
        return s > MAX_SLOTS ? MAX_SLOTS : s;
    }
};

__device__ __forceinline__ void group_sync(int group)
{
    asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(GROUP_THREADS) : "memory");
}

template <int MB, int NG>
__global__ void __launch_bounds__(prune_threads(NG), 1) prune_kernel(const PruneParams p)
{
    using L = PruneSmem<MB, NG>;
    constexpr int PFT = L::PFT;
    constexpr int CONSUMERS = NG * GROUP_WARPS;
    constexpr int NR = L::NR;
    constexpr int LDV = L::LDV;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* ring = reinterpret_cast<double*>(smem_raw);
    const int ring_bytes = L::ring_bytes(p.n_stages * CPS);
    const uint32_t stage_mask = (uint32_t)p.n_stages - 1u;
    double* slots = reinterpret_cast<double*>(smem_raw + ring_bytes);
    uint16_t* cnt_s = reinterpret_cast<uint16_t*>(smem_raw + ring_bytes + p.n_slots * L::SLOT_BYTES);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + ring_bytes + p.n_slots * L::SLOT_BYTES + PRUNE_CNT_CAP_BYTES);
    uint64_t* empty_bar = full_bar + MAX_STAGES;
    int* slot_exp = reinterpret_cast<int*>(smem_raw + ring_bytes + p.n_slots * L::SLOT_BYTES + PRUNE_CNT_CAP_BYTES + L::MISC_BYTES);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < p.n_stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], CONSUMERS);          // a stage is released when every consumer warp has read it
        }
        fence_barrier_init();
    }
    __syncthreads();

    const int64_t n_items = p.n_tiles * p.n_categories;

    if (warp >= CONSUMERS) {
        // ===== producers: stream the matrix K-chunks of every GEMM op of every item into the ring, alternate chunks each =====
        const int which = warp - CONSUMERS;
        if (lane == 0) {
            uint32_t pos = 0;
            for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int cat = (int)(item / p.n_tiles);
                const POp* ops = p.ops + (size_t)cat * p.n_ops;
                for (int o = 0; o < p.n_ops; ++o) {
                    const int type = ops[o].type;
                    if (type != OP_GEMM_SET && type != OP_GEMM_MUL && type != OP_GEMM_SET_LEAF && type != OP_GEMM_MUL_LEAF) continue;
                    const double* src = p.mp + (size_t)ops[o].mat * p.mp_stride;
                    for (int ch = 0; ch < p.n_kchunks; ch += CPS, ++pos) {
                        if ((int)(pos % PRODUCER_WARPS) != which) continue;
                        const uint32_t stage = pos & stage_mask;
                        const uint32_t round = pos >> p.stage_shift;
                        const uint32_t bytes = (uint32_t)min(CPS, p.n_kchunks - ch) * L::CHUNK_BYTES;      // the last stage of a GEMM may be partial
                        mbar_wait(&empty_bar[stage], (round & 1) ^ 1);
                        mbar_arrive_expect_tx(&full_bar[stage], bytes);
                        bulk_copy_g2s(ring + (size_t)stage * L::STAGE_DOUBLES, src + (size_t)ch * L::CHUNK_DOUBLES, bytes, &full_bar[stage]);
                    }
                }
            }
        }
        return;
    }

    // =============================== consumers =====================================================
    const int group = warp / GROUP_WARPS;         // families [group*16, +16) of the tile
    const int wg = warp % GROUP_WARPS;            // row group: rows [wg*8*MB, +8*MB)
    const int gtid = tid - group * GROUP_THREADS; // thread index inside the group
    const int g = lane >> 2;                      // fragment row / column group
    const int t4 = lane & 3;
    const int fbase = group * GFT;                // first tile family of this group
    uint32_t pos = 0;
    uint16_t* my_cnt = cnt_s + (size_t)fbase * p.n_leaves;

    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int cat = (int)(item / p.n_tiles);
        const int64_t tile = item % p.n_tiles;
        const int64_t fam0 = tile * PFT + fbase;  // first family of this group

        group_sync(group);      // previous item fully finished with this group's shared memory
        if (p.counts_in_smem) {
            const int total = GFT * p.n_leaves;
            for (int i = gtid; i < total; i += GROUP_THREADS) {
                const int f = i / p.n_leaves;
                int64_t fam = fam0 + f;
                if (fam >= p.n_families) fam = p.n_families - 1;
                my_cnt[i] = (uint16_t)p.counts[fam * p.n_leaves + (i - f * p.n_leaves)];
            }
        }
        if (gtid < MAX_SLOTS * GFT) slot_exp[(gtid / GFT) * PFT + fbase + (gtid % GFT)] = 0;
        group_sync(group);

        // the (L2-resident, 32-byte) ops are fetched one ahead: the load latency hides behind the current op
        const POp* gops = p.ops + (size_t)cat * p.n_ops;
        POp next_op = gops[0];
        for (int o = 0; o < p.n_ops; ++o) {
            const POp op = next_op;
            if (o + 1 < p.n_ops) next_op = gops[o + 1];
            switch (op.type) {
            case OP_LEAF_SET2: {
                // ---- a cherry in one pass: V = column(leaf 1) * column(leaf 2) ----
                const double* mt1 = p.mt + (size_t)op.mat * p.mt_stride;
                const double* mt2 = p.mt + (size_t)op.mat2 * p.mt_stride;
                double* dst = slots + (size_t)op.a * L::SLOT_DOUBLES;
                double v1[FPW][MB], v2[FPW][MB];
                #pragma unroll
                for (int fi = 0; fi < FPW; ++fi) {
                    const int f = wg * FPW + fi;
                    int o1, o2;
                    if (p.counts_in_smem) { o1 = my_cnt[f * p.n_leaves + op.col]; o2 = my_cnt[f * p.n_leaves + op.col2]; }
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
                for (int fi = 0; fi < FPW; ++fi) {
                    const int f = fbase + wg * FPW + fi;
                    double* row = dst + (size_t)f * LDV;
                    #pragma unroll
                    for (int i = 0; i < MB; ++i) row[lane + 32 * i] = v1[fi][i] * v2[fi][i];
                    if (lane == 0) slot_exp[op.a * PFT + f] = 0;
                }
                group_sync(group);
                break;
            }
            case OP_LEAF_SET:
            case OP_LEAF_MUL: {
                // ---- leaf edge: gather column obs (or an error-model stencil of columns) of M^T ----
                const double* mt = p.mt + (size_t)op.mat * p.mt_stride;
                const int col = op.col;
                double* dst = slots + (size_t)op.a * L::SLOT_DOUBLES;
                #pragma unroll
                for (int fi = 0; fi < FPW; ++fi) {
                    const int fl = wg * FPW + fi;
                    int obs;
                    if (p.counts_in_smem) obs = my_cnt[fl * p.n_leaves + col];
                    else {
                        int64_t fam = fam0 + fl;
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
                    const int f = fbase + fl;
                    double* row = dst + (size_t)f * LDV;
                    if (op.type == OP_LEAF_SET) {
                        #pragma unroll
                        for (int i = 0; i < MB; ++i) row[lane + 32 * i] = v[i];
                        if (lane == 0) slot_exp[op.a * PFT + f] = 0;     // fresh vector in a recycled slot
                    }
                    else {
                        #pragma unroll
                        for (int i = 0; i < MB; ++i) row[lane + 32 * i] *= v[i];
                    }
                }
                group_sync(group);
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
                const double* vsrc = slots + (size_t)src_slot * L::SLOT_DOUBLES + (size_t)(fbase + g) * LDV + t4;
                double acc[MB][2][2];
                #pragma unroll
                for (int i = 0; i < MB; ++i) acc[i][0][0] = acc[i][0][1] = acc[i][1][0] = acc[i][1][1] = 0.0;
                // Fragments are double-buffered in registers: the loads of panel q+1 (possibly from the next
                // ring stage) are issued before the MMAs of panel q, so shared-memory latency never gates the pipe.
                // A ring stage holds CPS chunks: the full / empty mbarriers are touched once per CPS * 20 DMMAs.
                const int a_off = (wg * 8 * MB) * 4 + lane;
                double a0[MB], a1[MB], b00, b01, b10, b11;
                uint32_t stage = pos & stage_mask;
                mbar_wait(&full_bar[stage], (pos >> p.stage_shift) & 1);
                {
                    const double* a_stage = ring + (size_t)stage * L::STAGE_DOUBLES + a_off;
                    #pragma unroll
                    for (int i = 0; i < MB; ++i) a0[i] = a_stage[i * 32];
                    b00 = vsrc[0];
                    b01 = vsrc[8 * LDV];
                }
                #pragma unroll 1
                for (int ch = 0; ch < p.n_kchunks; ++ch) {
                    const int sub = ch & (CPS - 1);                                     // chunk inside the stage
                    const double* a_chunk = ring + (size_t)stage * L::STAGE_DOUBLES + (size_t)sub * L::CHUNK_DOUBLES + a_off;
                    const int kcol = ch * (PPS * 4);
                    const bool last_of_stage = (sub == CPS - 1) || (ch + 1 == p.n_kchunks);
                    // panel 1 of this chunk
                    #pragma unroll
                    for (int i = 0; i < MB; ++i) a1[i] = a_chunk[NR * 4 + i * 32];
                    b10 = vsrc[kcol + 4];
                    b11 = vsrc[8 * LDV + kcol + 4];
                    #pragma unroll
                    for (int i = 0; i < MB; ++i) {
                        dmma_m8n8k4(acc[i][0][0], acc[i][0][1], a0[i], b00);
                        dmma_m8n8k4(acc[i][1][0], acc[i][1][1], a0[i], b01);
                    }
                    // panel 0 of the next chunk: same stage, or the next one (wait for it first)
                    const uint32_t npos = pos + (last_of_stage ? 1u : 0u);
                    const uint32_t nstage = npos & stage_mask;
                    if (ch + 1 < p.n_kchunks) {
                        if (last_of_stage) mbar_wait(&full_bar[nstage], (npos >> p.stage_shift) & 1);
                        const double* n_chunk = last_of_stage ? ring + (size_t)nstage * L::STAGE_DOUBLES + a_off : a_chunk + L::CHUNK_DOUBLES;
                        #pragma unroll
                        for (int i = 0; i < MB; ++i) a0[i] = n_chunk[i * 32];
                        b00 = vsrc[kcol + 8];
                        b01 = vsrc[8 * LDV + kcol + 8];
                    }
                    #pragma unroll
                    for (int i = 0; i < MB; ++i) {
                        dmma_m8n8k4(acc[i][0][0], acc[i][0][1], a1[i], b10);
                        dmma_m8n8k4(acc[i][1][0], acc[i][1][1], a1[i], b11);
                    }
                    if (last_of_stage) {
                        // every load of this stage has been consumed by an MMA above
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&empty_bar[stage]);
                    }
                    stage = nstage;
                    pos = npos;
                }
                // ---- epilogue, one FP64 multiply per element at most (none for a plain first factor): acc * 1.0 == acc
                //      exactly, so skipping the unit leaf factor changes no bit, and FP64 multiplies share the pipe
                //      with the other group's MMAs
                double* dst = slots + (size_t)op.a * L::SLOT_DOUBLES + (size_t)(fbase + t4 * 2) * LDV + wg * 8 * MB + g;
                if (with_leaf) {
                    // leaf-sibling factors, gathered after the K loop (the fragment registers are free now)
                    double lf[MB][2][2];
                    const double* mt2 = p.mt + (size_t)op.mat2 * p.mt_stride + wg * 8 * MB + g;
                    #pragma unroll
                    for (int nb = 0; nb < 2; ++nb)
                        #pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int fl = nb * 8 + t4 * 2 + e;
                            int obs;
                            if (p.counts_in_smem) obs = my_cnt[fl * p.n_leaves + op.col2];
                            else {
                                int64_t fam = fam0 + fl;
                                if (fam >= p.n_families) fam = p.n_families - 1;
                                obs = p.counts[fam * p.n_leaves + op.col2];
                            }
                            #pragma unroll
                            for (int i = 0; i < MB; ++i) lf[i][nb][e] = __ldg(mt2 + (size_t)obs * NR + i * 8);
                        }
                    if (is_set) group_sync(group);      // in place: every warp of the group is done reading V_child before anyone overwrites it
                    #pragma unroll
                    for (int i = 0; i < MB; ++i)
                        #pragma unroll
                        for (int nb = 0; nb < 2; ++nb) {
                            double* q0 = dst + (size_t)(nb * 8) * LDV + i * 8;
                            const double y0 = acc[i][nb][0] * lf[i][nb][0];
                            const double y1 = acc[i][nb][1] * lf[i][nb][1];
                            if (is_set) { q0[0] = y0; q0[LDV] = y1; }
                            else { q0[0] *= y0; q0[LDV] *= y1; }
                        }
                }
                else if (is_set) {
                    group_sync(group);                  // in place, as above
                    #pragma unroll
                    for (int i = 0; i < MB; ++i)
                        #pragma unroll
                        for (int nb = 0; nb < 2; ++nb) {
                            double* q0 = dst + (size_t)(nb * 8) * LDV + i * 8;
                            q0[0] = acc[i][nb][0];
                            q0[LDV] = acc[i][nb][1];
                        }
                }
                else {
                    double old[MB][2][2];               // all loads first, then the multiplies back to back
                    #pragma unroll
                    for (int i = 0; i < MB; ++i)
                        #pragma unroll
                        for (int nb = 0; nb < 2; ++nb) {
                            const double* q0 = dst + (size_t)(nb * 8) * LDV + i * 8;
                            old[i][nb][0] = q0[0];
                            old[i][nb][1] = q0[LDV];
                        }
                    #pragma unroll
                    for (int i = 0; i < MB; ++i)
                        #pragma unroll
                        for (int nb = 0; nb < 2; ++nb) {
                            double* q0 = dst + (size_t)(nb * 8) * LDV + i * 8;
                            q0[0] = old[i][nb][0] * acc[i][nb][0];
                            q0[LDV] = old[i][nb][1] * acc[i][nb][1];
                        }
                }
                if (!is_set && p.rescale && gtid < GFT) slot_exp[op.a * PFT + fbase + gtid] += slot_exp[op.b * PFT + fbase + gtid];
                group_sync(group);
                break;
            }
            case OP_SPILL:
            case OP_FILL: {
                // this group's half of the slot (families fbase .. fbase+GFT-1 are contiguous rows)
                double* sl = slots + (size_t)op.a * L::SLOT_DOUBLES + (size_t)fbase * LDV;
                double* sc = p.scratch + ((size_t)blockIdx.x * p.n_spill + op.b) * L::SLOT_DOUBLES + (size_t)fbase * LDV;
                int* sce = p.scratch_exp + ((size_t)blockIdx.x * p.n_spill + op.b) * PFT + fbase;
                if (op.type == OP_SPILL) {
                    for (int i = gtid; i < GFT * LDV; i += GROUP_THREADS) sc[i] = sl[i];
                    if (gtid < GFT) sce[gtid] = slot_exp[op.a * PFT + fbase + gtid];
                }
                else {
                    for (int i = gtid; i < GFT * LDV; i += GROUP_THREADS) sl[i] = sc[i];
                    if (gtid < GFT) slot_exp[op.a * PFT + fbase + gtid] = sce[gtid];
                }
                group_sync(group);
                break;
            }
            case OP_RESCALE: {
                // exact power-of-two renormalisation per family: V *= 2^-e, exponent tracked in slot_exp
                if (!p.rescale) break;          // uniform: reference arithmetic, no renormalisation
                double* sl = slots + (size_t)op.a * L::SLOT_DOUBLES;
                #pragma unroll
                for (int fi = 0; fi < FPW; ++fi) {
                    const int f = fbase + wg * FPW + fi;
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
                        if (lane == 0) slot_exp[op.a * PFT + f] += e;
                    }
                }
                group_sync(group);
                break;
            }
            case OP_ROOT: {
                // ---- root: index j <-> root size j+1 (src/base_model.cpp:95-98) ----
                const double* sl = slots + (size_t)op.a * L::SLOT_DOUBLES;
                #pragma unroll
                for (int fi = 0; fi < FPW; ++fi) {
                    const int fl = wg * FPW + fi;
                    const int f = fbase + fl;
                    const int64_t fam = fam0 + fl;
                    if (fam >= p.n_families) continue;      // warp-uniform
                    const double* row = sl + (size_t)f * LDV;
                    const int e = p.rescale ? slot_exp[op.a * PFT + f] : 0;
                    if (p.root_out) {
                        double* out = p.root_out + ((size_t)fam * p.n_categories + cat) * p.mrf;
                        for (int j = lane; j < p.mrf; j += 32) out[j] = e ? ldexp(row[j + 1], e) : row[j + 1];
                    }
                    if (p.mode == 2) {
                        // max_j L[j], no prior: the statistic of get_random_probabilities / compute_tree_pvalue
                        // (src/probability.cpp:308, 399)
                        double best = 0.0;
                        for (int j = lane; j < p.mrf; j += 32) {
                            double lj = row[j + 1];
                            if (e) lj = ldexp(lj, e);
                            best = fmax(best, lj);
                        }
                        #pragma unroll
                        for (int off = 16; off > 0; off >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, off));
                        if (lane == 0) { p.cat_lk[fam] = best; p.fail[fam] = 0; }
                    }
                    else if (p.mode == 0) {
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
