// Kernel 2+3 — fused Felsenstein-style pruning of a tile of NG x 16 families through the whole species
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
// Design (B200).  One persistent thread block per SM walks a host-built post-order PROGRAM for a tile of NG x 16
// families of one rate category.  The program is a stack machine (cafe_b200.cu, ProgramBuilder):
//
//   LEAVES   S' = prod of leaf columns                     a node whose children are all leaves (a cherry)
//   GEMM     acc = M_edge[NR x K] * S[K x 16]              an internal child c of node v, on the FP64 tensor pipe
//            acc = parked(v) * acc * leaf siblings         (the factors in the reference's Newick order)
//            -> parked(v) = acc                            if v has further internal children: PARK in tensor memory
//            -> S' = acc                                   otherwise v is complete: it is the next GEMM's source
//   ROOT     prior / category weight / max over root sizes of S
//
// Only ONE partial-likelihood vector per group has to be visible to all its warps: the source S of the next GEMM
// (B fragments).  It lives in shared memory, double-buffered (S, S'), so a finished vector is written while slower
// warps still read the old one: one group barrier per GEMM.  Every other live vector is a partial product waiting for
// a sibling subtree; it is only ever touched element-wise by the thread that owns the accumulator fragment, so it is
// PARKED IN TENSOR MEMORY (tcgen05.st / tcgen05.ld, 32x32b: each warp reaches exactly its own lane quarter, which is
// all an element-wise stack needs) — 256 KB per SM that an FP64 kernel would otherwise leave idle (tcgen05.mma has no
// FP64 kind; the contraction itself is mma.sync.m8n8k4.f64 = SASS DMMA).  Stack entries beyond the TMEM capacity go to
// an L2-resident scratch area (only trees whose Strahler-like depth exceeds the capacity).
//
// With the vector slots gone from shared memory (round 1: 3 slots x 32 families = 126 KB for two groups) there is room
// for THREE consumer groups (48-family tiles) next to an 80 KB matrix ring at N <= 160, which is what the FP64 pipe
// needs: a sub-partition's DMMA pipe only stays full while at least two of its warps are inside a K loop, and a group
// spends ~25 % of its time between K loops (epilogue, leaf gathers at L2 latency, barriers).
//
// The matrix streams L2 -> shared through a ring of stages (CPS K-chunks of 8 columns each) filled by 1-D bulk
// async copies (TMA engine, cp.async.bulk + mbarrier complete_tx) issued by producer warps that run ahead across ops
// and items; every group consumes the SAME stream (a stage is released when every consumer warp has read it), so the
// L2 traffic per flop is that of the whole tile.
//
// Leaf edges are gathers of one matrix column (or an error-model stencil of columns) — never GEMMs — done by the
// thread that owns the accumulator element; HBM traffic per family is its leaf counts in (1 byte each) and k+1 doubles out.
#pragma once

#include "common.cuh"

namespace cafe {

constexpr int GFT = 16;                                  // families per group = 2 n8 blocks
constexpr int MAX_GROUPS = 3;
constexpr int MAX_RING_STAGES = 16;

// Geometry of one instantiation, as plain functions so that the host plans shared memory with the same arithmetic.
// rb = 8-row blocks per warp, gw = warps per group (4: one per SM sub-partition; 8 for matrices above 256 rows),
// ng = groups, cps = K chunks per ring stage, pw = producer warps.
struct PruneGeom {
    int rb, gw, ng, cps, pw;
};
// Threads: the consumer warpgroups (one per group of four warps) plus ONE producer warpgroup, of which pw warps issue
// copies.  The register file is split per SM sub-partition (16 K registers each, one warp of every warpgroup), so the
// block is launched with 65536 / threads registers per thread and then re-balanced with setmaxnreg: the producer
// warpgroup drops to PRODUCER_REGS, every consumer warpgroup rises to pg_consumer_regs (wc * C + P <= 512 per lane).
constexpr int PRODUCER_REGS = 24;
__host__ __device__ constexpr int pg_consumer_wgs(int ng, int gw) { return ng * gw / 4; }
__host__ __device__ constexpr int pg_threads(int ng, int gw, int pw) { return (pg_consumer_wgs(ng, gw) + 1) * 128 + 0 * pw; }
__host__ __device__ constexpr int pg_consumer_regs(int ng, int gw)
{
    return (512 - PRODUCER_REGS) / pg_consumer_wgs(ng, gw) / 8 * 8 > 232 ? 232 : (512 - PRODUCER_REGS) / pg_consumer_wgs(ng, gw) / 8 * 8;
}
__host__ __device__ constexpr int pg_nr(int rb, int gw) { return 8 * rb * gw; }                  // padded matrix rows
__host__ __device__ constexpr int pg_ldv(int rb, int gw) { return pg_nr(rb, gw) + 4; }           // vector stride: % 16 == 4 -> conflict-free B fragments
__host__ __device__ constexpr int pg_chunk_bytes(int rb, int gw) { return PPS * 4 * pg_nr(rb, gw) * 8; }      // one K chunk: 8 matrix columns, panelised
__host__ __device__ constexpr int pg_stage_bytes(int rb, int gw, int cps) { return cps * pg_chunk_bytes(rb, gw); }
__host__ __device__ constexpr int pg_vec_bytes(int rb, int gw) { return GFT * pg_ldv(rb, gw) * 8; }           // one group's vector buffer
__host__ __device__ constexpr int pg_frag_cols(int rb) { return 8 * rb; }                       // one parked entry per thread, in 32-bit TMEM columns
__host__ __device__ constexpr int pg_tmem_capacity(int rb, int gw, int ng) { return 512 / (ng * (gw / 4) * pg_frag_cols(rb)); }      // parked entries per warp that fit
__host__ __device__ constexpr int pg_bar_bytes() { return 2 * MAX_RING_STAGES * 8 + 64; }
__host__ __device__ constexpr int pg_exp_bytes(int ng, int depth) { return ng * (2 + depth) * GFT * 4; }
__host__ __device__ constexpr int pg_total_bytes(int rb, int gw, int ng, int cps, int stages, int cnt_bytes, int depth)
{
    return stages * pg_stage_bytes(rb, gw, cps) + ng * 2 * pg_vec_bytes(rb, gw) + ((cnt_bytes + 15) & ~15) + pg_bar_bytes() + pg_exp_bytes(ng, depth);
}

template <int RB, int GW, int NG, int CPS, int PW>
struct PruneCfg {
    static constexpr int NR = pg_nr(RB, GW);
    static constexpr int LDV = pg_ldv(RB, GW);
    static constexpr int PFT = NG * GFT;                 // families per thread-block tile
    static constexpr int CONSUMERS = NG * GW;            // consumer warps
    static constexpr int THREADS = pg_threads(NG, GW, PW);
    static constexpr int CHUNK_BYTES = pg_chunk_bytes(RB, GW);
    static constexpr int CHUNK_DOUBLES = CHUNK_BYTES / 8;
    static constexpr int STAGE_BYTES = pg_stage_bytes(RB, GW, CPS);
    static constexpr int STAGE_DOUBLES = STAGE_BYTES / 8;
    static constexpr int VEC_BYTES = pg_vec_bytes(RB, GW);
    static constexpr int VEC_DOUBLES = VEC_BYTES / 8;
    static constexpr int FRAG_DOUBLES = 4 * RB;          // accumulator doubles per thread
    static constexpr int FRAG_COLS = pg_frag_cols(RB);
    static constexpr int BAR_BYTES = pg_bar_bytes();
    __host__ __device__ static constexpr int vec_bytes() { return NG * 2 * VEC_BYTES; }
};

__device__ __forceinline__ void group_sync(int group, int threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(threads) : "memory");
}

// ---- tensor memory: allocation and the 32x32b element-wise accessors (4 doubles = 8 columns per instruction) ----
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

__device__ __forceinline__ void tmem_st4(uint32_t taddr, double v0, double v1, double v2, double v3)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(__double2loint(v0)),
                 "r"(__double2hiint(v0)), "r"(__double2loint(v1)), "r"(__double2hiint(v1)), "r"(__double2loint(v2)), "r"(__double2hiint(v2)),
                 "r"(__double2loint(v3)), "r"(__double2hiint(v3))
                 : "memory");
}

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}

__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Leaf count i of a tile (1 or 2 bytes each; `base` is the staged tile in shared memory or the tile's rows in global memory).
__device__ __forceinline__ int load_count(const unsigned char* base, int i, int width)
{
    return width == 1 ? (int)base[i] : (int)reinterpret_cast<const uint16_t*>(base)[i];
}

// Leaf factor of ONE family column (RB values, rows row0 + STEP i: STEP = 8 in the accumulator-fragment layout, 32 in the
// row-major passes): column `obs` of the
// leaf's transposed matrix, or — with an error model — sum over deviations, ascending child size, no FMA
// (src/probability.cpp:182-193 feeding src/matrix_cache.cpp:48-54).  mt points at row0 of the leaf's matrix.
template <int RB, int NR, int STEP = 8>
__device__ __forceinline__ void leaf_column(double (&v)[RB], const double* mt, int obs, const double* err, int err_ndev, int mf)
{
    if (err == nullptr) {
        #pragma unroll
        for (int i = 0; i < RB; ++i) v[i] = __ldg(mt + (size_t)obs * NR + i * STEP);
    }
    else {
        #pragma unroll
        for (int i = 0; i < RB; ++i) v[i] = 0.0;
        const int offset = obs - (err_ndev - 1) / 2;
        for (int d = 0; d < err_ndev; ++d) {
            const int c = offset + d;
            if (c < 0 || c > mf) continue;
            const double pe = __ldg(err + (size_t)obs * err_ndev + d);
            #pragma unroll
            for (int i = 0; i < RB; ++i) v[i] = __dadd_rn(v[i], __dmul_rn(__ldg(mt + (size_t)c * NR + i * STEP), pe));
        }
    }
}

template <int RB, int GW, int NG, int CPS, int PW>
__global__ void __launch_bounds__(pg_threads(NG, GW, PW), 1) prune_kernel(const PruneParams p)
{
    using L = PruneCfg<RB, GW, NG, CPS, PW>;
    constexpr int NR = L::NR;
    constexpr int LDV = L::LDV;
    constexpr int PFT = L::PFT;
    constexpr int CONSUMERS = L::CONSUMERS;
    constexpr int GROUP_THREADS = GW * 32;
    constexpr int FPW = GFT / GW;                 // families per warp in the row-major ops (leaf products, root)
    constexpr int RPL = NR / 32;                  // rows per lane in the row-major ops
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* ring = reinterpret_cast<double*>(smem_raw);
    const int ring_bytes = p.n_stages * L::STAGE_BYTES;
    double* vecs = reinterpret_cast<double*>(smem_raw + ring_bytes);
    unsigned char* cnt_s = smem_raw + ring_bytes + L::vec_bytes();                  // staged leaf counts of the tile (1 or 2 bytes each)
    unsigned char* after_cnt = smem_raw + ring_bytes + L::vec_bytes() + ((p.cnt_smem_bytes + 15) & ~15);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(after_cnt);
    uint64_t* empty_bar = full_bar + MAX_RING_STAGES;
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(empty_bar + MAX_RING_STAGES);
    int* exps = reinterpret_cast<int*>(after_cnt + L::BAR_BYTES);        // [NG][2 + depth][GFT]: S buffers, then the parked stack
    // the program of the current rate category, when it fits: [n_ops] POp, then [n_leafrefs] LeafRef
    const POp* ops_s = reinterpret_cast<const POp*>(after_cnt + L::BAR_BYTES + ((pg_exp_bytes(NG, p.depth) + 15) & ~15));
    const LeafRef* leaf_s = reinterpret_cast<const LeafRef*>(ops_s + p.n_ops);

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
    if (warp == 0 && p.tmem_cols > 0) tmem_alloc(tmem_base_s, (uint32_t)p.tmem_cols);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    const int64_t n_items = p.n_tiles * p.n_categories;

    if (warp >= CONSUMERS) {
        // ===== producer warpgroup: give registers back, then PW of its warps stream the matrix K-chunks of every GEMM op
        //       of every item into the ring, alternate stages each =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PRODUCER_REGS));
        const int which = warp - CONSUMERS;
        if (which < PW && lane == 0) {
            uint32_t stage = 0, phase = 0, turn = 0;
            for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int cat = (int)(item / p.n_tiles);
                const POp* ops = p.ops + (size_t)cat * p.n_ops;
                for (int o = 0; o < p.n_ops; ++o) {
                    const int2 tm = *reinterpret_cast<const int2*>(&ops[o]);        // {type, mat}
                    if (tm.x != POP_GEMM) continue;
                    const double* src = p.mp + (size_t)tm.y * p.mp_stride;
                    for (int ch = 0; ch < p.n_kchunks; ch += CPS) {
                        if (PW == 1 || (int)turn == which) {
                            const uint32_t bytes = (uint32_t)min(CPS, p.n_kchunks - ch) * L::CHUNK_BYTES;      // the last stage of a GEMM may be partial
                            mbar_wait(&empty_bar[stage], phase ^ 1);
                            mbar_arrive_expect_tx(&full_bar[stage], bytes);
                            bulk_copy_g2s(ring + (size_t)stage * L::STAGE_DOUBLES, src + (size_t)ch * L::CHUNK_DOUBLES, bytes, &full_bar[stage]);
                        }
                        if (PW > 1) turn = (turn + 1 == PW) ? 0 : turn + 1;
                        if (++stage == (uint32_t)p.n_stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
        return;
    }

    // =============================== consumers =====================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(pg_consumer_regs(NG, GW)));
    const int group = warp / GW;                  // families [group*16, +16) of the tile
    const int wg = warp % GW;                     // row group: rows [wg*8*RB, +8*RB)
    const int gtid = tid - group * GROUP_THREADS; // thread index inside the group
    const int g = lane >> 2;                      // fragment row / column group
    const int t4 = lane & 3;
    const int fbase = group * GFT;                // first tile family of this group
    const int row0 = wg * 8 * RB + g;             // this thread's first accumulator row
    uint32_t stage = 0, phase = 0;
    double* gvec = vecs + (size_t)group * 2 * L::VEC_DOUBLES;          // this group's two vector buffers
    // 32-bit shared addresses of the K loop's operands: this thread's A fragment (row block 0) in ring stage 0, its B
    // fragment in vector buffer 0, the mbarrier arrays
    // (pinned in registers: the compiler would otherwise re-derive them from %tid and the shared window at every stage)
    const uint32_t ring_u = pin_u32(smem_u32(ring) + (uint32_t)((wg * 8 * RB) * 4 + lane) * 8u);
    const uint32_t vec_u = pin_u32(smem_u32(gvec) + (uint32_t)(g * LDV + t4) * 8u);
    const uint32_t full_u = pin_u32(smem_u32(full_bar)), empty_u = pin_u32(smem_u32(empty_bar));
    int* gexp = exps + group * (2 + p.depth) * GFT;
    // parked entries of this warp: TMEM lane quarter = warp % 4, columns by (sharer, entry)
    const uint32_t tmem_warp = (p.tmem_cols > 0 ? *tmem_base_s : 0u) + ((uint32_t)((warp & 3) * 32) << 16) +
                               (uint32_t)((group * (GW / 4) + wg / 4) * p.tmem_entries * L::FRAG_COLS);
    double* gscratch = p.scratch + (size_t)blockIdx.x * p.n_gspill * (CONSUMERS * 32) * L::FRAG_DOUBLES + (size_t)(warp * 32 + lane);
    int cur = 0;                                  // which of the two buffers holds the current complete vector
    int loaded_cat = -1;                          // category whose program sits in shared memory

    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int cat = (int)(item / p.n_tiles);
        const int64_t tile = item - (int64_t)cat * p.n_tiles;
        const int64_t fam0 = tile * PFT + fbase;  // first family of this group (the count matrix is padded to whole tiles)
        // this group's 16 count rows: the staged tile, or straight from the (padded) global matrix
        const unsigned char* cnt = p.counts_in_smem ? cnt_s + (size_t)fbase * p.n_leaves * p.cnt_width
                                                    : reinterpret_cast<const unsigned char*>(p.counts) + fam0 * p.n_leaves * p.cnt_width;

        if (p.ops_in_smem && cat != loaded_cat) {
            // category change (at most n_categories times per block): every consumer warp has left the old program
            asm volatile("bar.sync 14, %0;" ::"n"(CONSUMERS * 32) : "memory");
            const int n16 = (p.n_ops * (int)sizeof(POp) + p.n_leafrefs * (int)sizeof(LeafRef) + 15) / 16;
            const int n16_ops = p.n_ops * (int)sizeof(POp) / 16;
            const uint4* src_ops = reinterpret_cast<const uint4*>(p.ops + (size_t)cat * p.n_ops);
            const uint4* src_leaf = reinterpret_cast<const uint4*>(p.leaves + (size_t)cat * p.n_leafrefs);      // 16-byte aligned: n_leafrefs is even
            uint4* dst = reinterpret_cast<uint4*>(const_cast<POp*>(ops_s));
            for (int i = tid; i < n16; i += CONSUMERS * 32) dst[i] = i < n16_ops ? __ldg(src_ops + i) : __ldg(src_leaf + (i - n16_ops));
            asm volatile("bar.sync 14, %0;" ::"n"(CONSUMERS * 32) : "memory");
            loaded_cat = cat;
        }
        group_sync(group, GROUP_THREADS);         // previous item fully finished with this group's shared memory
        if (p.counts_in_smem) {
            // the group's 16 count rows are one contiguous, 16-byte aligned block of the row-major matrix
            const int n16 = GFT * p.n_leaves * p.cnt_width / 16;
            const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(p.counts) + fam0 * p.n_leaves * p.cnt_width);
            uint4* dst = reinterpret_cast<uint4*>(cnt_s + (size_t)fbase * p.n_leaves * p.cnt_width);
            for (int i = gtid; i < n16; i += GROUP_THREADS) dst[i] = __ldg(src + i);
        }
        if (p.rescale && gtid < 2 * GFT) gexp[gtid] = 0;
        group_sync(group, GROUP_THREADS);

        // the 32-byte ops come from shared memory (or, for programs too long for it, from L2)
        const POp* gops = p.ops_in_smem ? ops_s : p.ops + (size_t)cat * p.n_ops;
        const LeafRef* gleaves = p.ops_in_smem ? leaf_s : p.leaves + (size_t)cat * p.n_leafrefs;
        for (int o = 0; o < p.n_ops; ++o) {
            const int op_type = gops[o].type;
            if (op_type == POP_GEMM) {
                // ---- internal edge: acc = M * S on the FP64 tensor pipe ----
                // All shared-memory operands are addressed with 32-bit shared addresses computed once per GEMM and read with
                // ld.shared: the loop then carries two addresses and the ring position, nothing is re-derived per chunk.
                double acc[RB][2][2];
                #pragma unroll
                for (int i = 0; i < RB; ++i) acc[i][0][0] = acc[i][0][1] = acc[i][1][0] = acc[i][1][1] = 0.0;
                {
                    // Fragments are double-buffered in registers: the loads of panel q+1 (possibly from the next ring stage)
                    // are issued before the MMAs of panel q, so shared-memory latency never gates the pipe.  A ring stage holds
                    // CPS chunks: the full / empty mbarriers are touched once per CPS * 4 * RB DMMAs.
                    constexpr uint32_t PANEL_BYTES = NR * 4 * 8;                   // one K panel: 4 columns x NR rows
                    constexpr uint32_t NB1 = 8u * LDV * 8u;                        // second n8 block of the B operand
                    uint32_t vb = vec_u + (uint32_t)cur * L::VEC_BYTES;            // B fragments: V[g][k + t4], V[8 + g][k + t4]
                    double a0[RB], a1[RB], b00, b01, b10, b11;
                    mbar_wait_u32(full_u + stage * 8u, phase);
                    uint32_t sa = ring_u + stage * (uint32_t)L::STAGE_BYTES;       // this thread's A fragments of the current stage
                    #pragma unroll
                    for (int i = 0; i < RB; ++i) a0[i] = lds_f64(sa + i * 256u);
                    b00 = lds_f64(vb);
                    b01 = lds_f64(vb + NB1);
                    int left = p.n_kchunks;                                        // chunks of this GEMM not yet multiplied
                    #pragma unroll 1
                    while (left > 0) {
                        const int nc = left < CPS ? left : CPS;                    // chunks in this stage (the last stage may be partial)
                        uint32_t nstage = stage + 1, nphase = phase;
                        if (nstage == (uint32_t)p.n_stages) { nstage = 0; nphase ^= 1; }
                        #pragma unroll
                        for (int c = 0; c < CPS; ++c) {
                            if (c < nc) {
                                // panel 1 of this chunk
                                #pragma unroll
                                for (int i = 0; i < RB; ++i) a1[i] = lds_f64(sa + (2 * c + 1) * PANEL_BYTES + i * 256u);
                                b10 = lds_f64(vb + (2 * c + 1) * 32u);
                                b11 = lds_f64(vb + NB1 + (2 * c + 1) * 32u);
                                #pragma unroll
                                for (int i = 0; i < RB; ++i) {
                                    dmma_m8n8k4(acc[i][0][0], acc[i][0][1], a0[i], b00);
                                    dmma_m8n8k4(acc[i][1][0], acc[i][1][1], a0[i], b01);
                                }
                                // panel 0 of the next chunk: same stage, or the next one (wait for it first)
                                if (c + 1 < nc) {
                                    #pragma unroll
                                    for (int i = 0; i < RB; ++i) a0[i] = lds_f64(sa + (2 * c + 2) * PANEL_BYTES + i * 256u);
                                    b00 = lds_f64(vb + (2 * c + 2) * 32u);
                                    b01 = lds_f64(vb + NB1 + (2 * c + 2) * 32u);
                                }
                                else if (left > nc) {
                                    mbar_wait_u32(full_u + nstage * 8u, nphase);
                                    const uint32_t na = ring_u + nstage * (uint32_t)L::STAGE_BYTES;
                                    #pragma unroll
                                    for (int i = 0; i < RB; ++i) a0[i] = lds_f64(na + i * 256u);
                                    b00 = lds_f64(vb + (2 * c + 2) * 32u);
                                    b01 = lds_f64(vb + NB1 + (2 * c + 2) * 32u);
                                }
                                #pragma unroll
                                for (int i = 0; i < RB; ++i) {
                                    dmma_m8n8k4(acc[i][0][0], acc[i][0][1], a1[i], b10);
                                    dmma_m8n8k4(acc[i][1][0], acc[i][1][1], a1[i], b11);
                                }
                            }
                        }
                        // every load of this stage has been consumed by an MMA above
                        __syncwarp();
                        if (lane == 0) mbar_arrive_u32(empty_u + stage * 8u);
                        stage = nstage;
                        phase = nphase;
                        sa = ring_u + stage * (uint32_t)L::STAGE_BYTES;
                        vb += nc * 64u;
                        left -= nc;
                    }
                }

                // the op's other fields are read only now, so that nothing of them is live across the K loop
                const POp op = gops[o];
                // ---- epilogue in registers: the factors in the reference's order (src/probability.cpp:229-238):
                //      [leaves before] * (parked product | this GEMM) * leaves after.  a * b == b * a exactly, so
                //      "parked * acc" and "leaf * acc" are the reference's left-to-right products bit for bit.
                // The fragment is walked one family column (nb, e) at a time — RB values — so that the accumulators plus a
                // handful of temporaries are all that is live (three consumer groups leave 160 registers per thread).
                const LeafRef* lrs = gleaves + op.leaf_begin;
                if (op.n_pre > 0) {
                    // leaves before the first internal child (nodes with more than two children only): (l1 * l2 ...) * acc
                    #pragma unroll
                    for (int nb = 0; nb < 2; ++nb)
                        #pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int fl = nb * 8 + t4 * 2 + e;
                            double t[RB];
                            for (int q = 0; q < op.n_pre; ++q) {
                                const LeafRef lr = lrs[q];
                                double v[RB];
                                leaf_column<RB, NR>(v, p.mt + (size_t)lr.mat * p.mt_stride + row0, load_count(cnt, fl * p.n_leaves + lr.col, p.cnt_width),
                                                    p.err, p.err_ndev, p.mf);
                                #pragma unroll
                                for (int i = 0; i < RB; ++i) t[i] = (q == 0) ? v[i] : __dmul_rn(t[i], v[i]);
                            }
                            #pragma unroll
                            for (int i = 0; i < RB; ++i) acc[i][nb][e] = __dmul_rn(t[i], acc[i][nb][e]);
                        }
                }
                if (op.flags & PF_PARKED) {
                    if (op.park >= 0) {
                        const uint32_t ta = tmem_warp + (uint32_t)(op.park * L::FRAG_COLS);
                        #pragma unroll
                        for (int i = 0; i < RB; ++i) {
                            uint32_t r[8];
                            tmem_ld4(ta + i * 8, r);
                            tmem_wait_ld();
                            #pragma unroll
                            for (int nb = 0; nb < 2; ++nb) {
                                acc[i][nb][0] = __dmul_rn(__hiloint2double(r[nb * 4 + 1], r[nb * 4 + 0]), acc[i][nb][0]);
                                acc[i][nb][1] = __dmul_rn(__hiloint2double(r[nb * 4 + 3], r[nb * 4 + 2]), acc[i][nb][1]);
                            }
                        }
                    }
                    else {
                        const double* sc = gscratch + (size_t)(-op.park - 1) * (CONSUMERS * 32) * L::FRAG_DOUBLES;
                        #pragma unroll
                        for (int i = 0; i < RB; ++i)
                            #pragma unroll
                            for (int nb = 0; nb < 2; ++nb) {
                                acc[i][nb][0] = __dmul_rn(sc[(size_t)((i * 2 + nb) * 2 + 0) * (CONSUMERS * 32)], acc[i][nb][0]);
                                acc[i][nb][1] = __dmul_rn(sc[(size_t)((i * 2 + nb) * 2 + 1) * (CONSUMERS * 32)], acc[i][nb][1]);
                            }
                    }
                }
                // leaves after the internal child: two L1 / L2 round trips per leaf — the loads of two family columns of the
                // fragment are in flight together (all four would not fit the 160 registers of the three-group geometry)
                for (int q = 0; q < op.n_post; ++q) {
                    const LeafRef lr = lrs[op.n_pre + q];
                    const double* mt = p.mt + (size_t)lr.mat * p.mt_stride + row0;
                    #pragma unroll
                    for (int nb = 0; nb < 2; ++nb) {
                        double v[2][RB];
                        #pragma unroll
                        for (int e = 0; e < 2; ++e)
                            leaf_column<RB, NR>(v[e], mt, load_count(cnt, (nb * 8 + t4 * 2 + e) * p.n_leaves + lr.col, p.cnt_width), p.err, p.err_ndev, p.mf);
                        #pragma unroll
                        for (int e = 0; e < 2; ++e)
                            #pragma unroll
                            for (int i = 0; i < RB; ++i) acc[i][nb][e] = __dmul_rn(acc[i][nb][e], v[e][i]);
                    }
                }
                // exponent bookkeeping of the optional power-of-two rescaling: the product carries the sum of its factors' exponents
                int* e_par = gexp + (2 + (op.park >= 0 ? op.park + p.n_gspill : -op.park - 1)) * GFT;     // stack entry -> exponent row
                if (p.rescale && gtid < GFT) {
                    int e = gexp[cur * GFT + gtid];
                    if (op.flags & PF_PARKED) e += e_par[gtid];
                    if (op.flags & PF_PARK) e_par[gtid] = e;
                    else gexp[(cur ^ 1) * GFT + gtid] = e;
                }
                if (op.flags & PF_PARK) {
                    // the parent waits for another internal child: park the partial product, nobody else needs to see it
                    if (op.park >= 0) {
                        const uint32_t ta = tmem_warp + (uint32_t)(op.park * L::FRAG_COLS);
                        #pragma unroll
                        for (int i = 0; i < RB; ++i) tmem_st4(ta + i * 8, acc[i][0][0], acc[i][0][1], acc[i][1][0], acc[i][1][1]);
                        tmem_wait_st();
                    }
                    else {
                        double* sc = gscratch + (size_t)(-op.park - 1) * (CONSUMERS * 32) * L::FRAG_DOUBLES;
                        #pragma unroll
                        for (int i = 0; i < RB; ++i)
                            #pragma unroll
                            for (int nb = 0; nb < 2; ++nb) {
                                sc[(size_t)((i * 2 + nb) * 2 + 0) * (CONSUMERS * 32)] = acc[i][nb][0];
                                sc[(size_t)((i * 2 + nb) * 2 + 1) * (CONSUMERS * 32)] = acc[i][nb][1];
                            }
                    }
                    // the source buffer stays "current" until the next complete vector is written into the other one
                }
                else {
                    // the parent is complete: it becomes the next GEMM's source, written into the idle buffer
                    double* dst = gvec + (size_t)(cur ^ 1) * L::VEC_DOUBLES + (size_t)(t4 * 2) * LDV + wg * 8 * RB + g;
                    #pragma unroll
                    for (int i = 0; i < RB; ++i)
                        #pragma unroll
                        for (int nb = 0; nb < 2; ++nb) {
                            double* q0 = dst + (size_t)(nb * 8) * LDV + i * 8;
                            q0[0] = acc[i][nb][0];
                            q0[LDV] = acc[i][nb][1];
                        }
                    cur ^= 1;
                    group_sync(group, GROUP_THREADS);
                }
            }
            else if (op_type == POP_LEAVES) {
                // ---- a node whose children are all leaves: S' = column(leaf 1) * column(leaf 2) * ... (row-major pass) ----
                const POp op = gops[o];
                double* dst = gvec + (size_t)(cur ^ 1) * L::VEC_DOUBLES;
                const LeafRef* lrs = gleaves + op.leaf_begin;
                // Two families of the warp per round trip, and the first two leaves (a cherry) in the same one: 4 * RPL loads
                // in flight instead of RPL at a time.
                constexpr int HF = FPW >= 2 ? 2 : 1;                 // families per batch
                #pragma unroll
                for (int f0 = 0; f0 < FPW; f0 += HF) {
                    double v[HF][RPL];
                    {
                        const LeafRef la = lrs[0];
                        const bool two = op.n_pre > 1;
                        const LeafRef lb = two ? lrs[1] : la;
                        double wa[HF][RPL], wb[HF][RPL];
                        #pragma unroll
                        for (int h = 0; h < HF; ++h) {
                            const int fl = wg * FPW + f0 + h;
                            leaf_column<RPL, NR, 32>(wa[h], p.mt + (size_t)la.mat * p.mt_stride + lane, load_count(cnt, fl * p.n_leaves + la.col, p.cnt_width),
                                                     p.err, p.err_ndev, p.mf);
                            if (two)
                                leaf_column<RPL, NR, 32>(wb[h], p.mt + (size_t)lb.mat * p.mt_stride + lane,
                                                         load_count(cnt, fl * p.n_leaves + lb.col, p.cnt_width), p.err, p.err_ndev, p.mf);
                        }
                        #pragma unroll
                        for (int h = 0; h < HF; ++h)
                            #pragma unroll
                            for (int i = 0; i < RPL; ++i) v[h][i] = two ? __dmul_rn(wa[h][i], wb[h][i]) : wa[h][i];
                    }
                    for (int q = 2; q < op.n_pre; ++q) {
                        const LeafRef lr = lrs[q];
                        double w[HF][RPL];
                        #pragma unroll
                        for (int h = 0; h < HF; ++h)
                            leaf_column<RPL, NR, 32>(w[h], p.mt + (size_t)lr.mat * p.mt_stride + lane,
                                                     load_count(cnt, (wg * FPW + f0 + h) * p.n_leaves + lr.col, p.cnt_width), p.err, p.err_ndev, p.mf);
                        #pragma unroll
                        for (int h = 0; h < HF; ++h)
                            #pragma unroll
                            for (int i = 0; i < RPL; ++i) v[h][i] = __dmul_rn(v[h][i], w[h][i]);
                    }
                    #pragma unroll
                    for (int h = 0; h < HF; ++h) {
                        double* row = dst + (size_t)(wg * FPW + f0 + h) * LDV + lane;
                        #pragma unroll
                        for (int i = 0; i < RPL; ++i) row[32 * i] = v[h][i];
                    }
                }
                if (p.rescale && gtid < GFT) gexp[(cur ^ 1) * GFT + gtid] = 0;
                cur ^= 1;
                group_sync(group, GROUP_THREADS);
            }
            else if (op_type == POP_ROOT) {
                // ---- root: index j <-> root size j+1 (src/base_model.cpp:95-98) ----
                const double* sl = gvec + (size_t)cur * L::VEC_DOUBLES;
                #pragma unroll
                for (int fi = 0; fi < FPW; ++fi) {
                    const int fl = wg * FPW + fi;
                    const int64_t fam = fam0 + fl;
                    if (fam >= p.n_families) continue;      // warp-uniform
                    const double* row = sl + (size_t)fl * LDV;
                    const int e = p.rescale ? gexp[cur * GFT + fl] : 0;
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
                            p.cat_lk[(size_t)cat * p.n_families + fam] = best * p.cat_probs[cat];
                            p.fail[(size_t)cat * p.n_families + fam] = nonzero ? 0 : 1;
                        }
                    }
                }
            }
            // ---- optional exact power-of-two renormalisation of the vector just completed: V *= 2^-e per family,
            //      exponent tracked beside the buffer.  Below the root only rows <= mf feed the next contraction; the rows
            //      above are zeroed so that a scaled-up dead entry can never meet a zero matrix column as inf * 0.
            const bool completed = (op_type == POP_LEAVES) || (op_type == POP_GEMM && !(gops[o].flags & PF_PARK));
            if (p.rescale && completed) {
                const bool root_vec = (o + 1 < p.n_ops) && gops[o + 1].type == POP_ROOT;
                const int top = root_vec ? max(p.mf, p.mrf) : p.mf;
                double* sl = gvec + (size_t)cur * L::VEC_DOUBLES;
                #pragma unroll
                for (int fi = 0; fi < FPW; ++fi) {
                    const int fl = wg * FPW + fi;
                    double* row = sl + (size_t)fl * LDV;
                    double m = 0.0;
                    for (int s = lane; s <= top; s += 32) m = fmax(m, row[s]);
                    #pragma unroll
                    for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
                    if (m > 0.0 && m < 0x1p-64) {
                        int e;
                        frexp(m, &e);                       // m = f * 2^e, f in [0.5, 1)
                        for (int s = lane; s < NR; s += 32) row[s] = (s <= top) ? ldexp(row[s], -e) : 0.0;     // exact, subnormals included
                        if (lane == 0) gexp[cur * GFT + fl] += e;
                    }
                }
                group_sync(group, GROUP_THREADS);
            }
        }
    }

    // all consumer warps are done with tensor memory
    if (p.tmem_cols > 0) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync 15, %0;" ::"n"(CONSUMERS * 32) : "memory");
        if (warp == 0) tmem_dealloc(*tmem_base_s, (uint32_t)p.tmem_cols);
    }
}

}  // namespace cafe
