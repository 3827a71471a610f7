// Shared definitions for the sm_100a kernels of the birth-death likelihood engine.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace cafe {

// ---------------------------------------------------------------------------------------------
// Geometry of the pruning kernel.  A thread block prunes FT families ("columns") of one rate
// category through the whole tree; every partial-likelihood vector lives in shared memory as
// V[family][size] with a padded stride so that FP64 MMA B-fragments load without bank conflicts.
// ---------------------------------------------------------------------------------------------
constexpr int FT = 32;                  // families per tile (MMA N dimension = 4 n8 blocks)
constexpr int STAGES = 4;               // ring of matrix K-chunks (reconstruction kernel)
constexpr int MAX_STAGES = 8;           // pruning kernel: 2, 4 or 8 stages chosen at create (power of two)
constexpr int PPS = 2;                  // K panels (of 4 columns) per stage
constexpr int CNT_CAP_BYTES = 8192;     // staged leaf counts (uint16) per tile, if they fit
constexpr int LGAMMA_TABLE = 1024;      // src/probability.cpp:52
constexpr int MAX_SLOTS = 8;
constexpr int MAX_MB = 8;               // rows padded to NR = 32*MB <= 256

__host__ __device__ constexpr int nr_of(int mb) { return 32 * mb; }
__host__ __device__ constexpr int ldv_of(int mb) { return 32 * mb + 4; }            // stride % 16 == 4 (8-byte words)
__host__ __device__ constexpr int stage_doubles(int mb) { return PPS * 4 * 32 * mb; }

// Slot-machine op codes of the reconstruction kernel's schedule (host-built, one list per tree; ScheduleBuilder in cafe_b200.cu)
enum : int {
    OP_LEAF_SET = 0,   // V[a] = column obs of M(node)            (leaf child, first factor of its parent)
    OP_LEAF_MUL = 1,   // V[a] *= column obs of M(node)
    OP_GEMM_SET = 2,   // V[a] = M(node) * V[a]                   (internal child, first factor)
    OP_GEMM_MUL = 3,   // V[a] *= M(node) * V[b]
    OP_SPILL = 4,      // scratch[b] = V[a]
    OP_FILL = 5,       // V[a] = scratch[b]
    OP_RESCALE = 6,    // (unused by the kernels; marks the end of an internal node)
    OP_ROOT = 7        // root prior, write results
};

// Tree-level op as the host schedule builder emits it (what the reconstruction kernel walks).
struct Op {
    int type;
    int a;
    int b;
    int node;
};

// Stack-machine program of the pruning kernel (ProgramBuilder in cafe_b200.cu; semantics in prune.cuh).
enum : int { POP_LEAVES = 0, POP_GEMM = 1, POP_ROOT = 2 };
constexpr int PF_PARKED = 1;   // GEMM: multiply by the partial product parked at `park` (pop)
constexpr int PF_PARK = 2;     // GEMM: park the result at `park` (push) instead of completing the parent's vector

// Pruning op with everything resolved for one rate category: one aligned 32-byte load per op.
struct __align__(16) POp {
    int type;
    int mat;         // GEMM: unique-matrix slot of the child's edge
    int flags;
    int park;        // stack entry: >= 0 tensor-memory entry, < 0: -(device scratch entry + 1)
    int leaf_begin;  // this op's leaf list in the per-category LeafRef array: n_pre entries, then n_post
    int n_pre;       // GEMM: leaves multiplied before the GEMM factor (n-ary nodes, Newick order); LEAVES: all of them
    int n_post;      // GEMM: leaves multiplied after it
    int node;
};

struct LeafRef {
    int mat;         // unique-matrix slot of the leaf's edge
    int col;         // its column of the count matrix
};

struct PruneParams {
    // problem
    int64_t n_families;
    int n_leaves;
    int n_nodes;
    int n_categories;
    int mf;                 // max_family_size
    int mrf;                // max_root_family_size
    int n_ops;
    int n_leafrefs;
    int n_kchunks;          // K chunks per GEMM (each PPS*4 columns)
    int mode;
    int rescale;
    int err_rows;
    int err_ndev;
    int counts_in_smem;
    int cnt_smem_bytes;     // staged leaf counts of one tile (0 when they do not fit)
    int cnt_width;          // bytes per leaf count on the device: 1 (max_family_size <= 255) or 2
    int ops_in_smem;        // the per-category program is staged in shared memory
    int n_stages;           // ring depth
    int depth;              // parked-stack depth of the program
    int n_gspill;           // ... of which the outermost entries live in device scratch
    int tmem_entries;       // ... and the rest in tensor memory (entries per warp)
    int tmem_cols;          // tensor-memory columns to allocate (power of two >= 32, 0 = none)
    int64_t n_tiles;        // tiles of (groups x 16) families = work items per category
    // device pointers
    const POp* ops;                 // [k][n_ops]
    const LeafRef* leaves;          // [k][n_leafrefs]
    const void* counts;             // [F padded to whole tiles][n_leaves] uint8 (max_family_size <= 255) or uint16
    const double* mp;               // panelised matrices   [U][kpanels][NR][4]
    const double* mt;               // transposed matrices  [U][mf+1][NR]
    size_t mp_stride;               // doubles per matrix
    size_t mt_stride;
    const double* err;              // [err_rows][err_ndev] or null
    const double* prior;            // [mrf]
    const double* logprior;         // [mrf]
    const double* cat_probs;        // [k]
    double* scratch;                // [grid][n_gspill][4*RB][consumer threads]
    // outputs
    double* cat_lk;                 // [k][F]   (gamma: category-major, so that a category pass writes whole sectors)  or family lnL [F] (base)
    uint8_t* fail;                  // [k][F]
    double* root_out;               // [F][k][mrf] or null (inspection)
};

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier, bulk async copy (TMA engine, 1-D), FP64 tensor-core MMA.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// The same on 32-bit shared addresses (the K loop keeps them in registers instead of re-deriving them from pointers).
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP_U:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE_U;\n"
        "bra WAIT_LOOP_U;\n"
        "WAIT_DONE_U:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void mbar_arrive_u32(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Opaque copy: the value lives in a register from here on (no rematerialisation from its inputs).
__device__ __forceinline__ uint32_t pin_u32(uint32_t x)
{
    uint32_t y;
    asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}

__device__ __forceinline__ double lds_f64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// 1-D bulk copy global -> shared through the TMA engine; completion is signalled on the mbarrier.
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// D(8x8) += A(8x4, row) * B(4x8, col) on the FP64 tensor pipe (SASS: DMMA).
// lane l holds A[l/4][l%4], B[l%4][l/4], D[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

}  // namespace cafe
