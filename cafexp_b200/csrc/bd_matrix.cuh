// Kernel 1 — birth-death transition matrices for every unique (branch length, lambda) key of an
// evaluation, all rate categories at once.
//
// Restates (file:line in the reference)
//   birthdeath_rate_with_log_alpha                       src/probability.cpp:101-145
//   the_probability_of_going_from_parent_fam_size_to_c   src/probability.cpp:147-164
//   matrix_cache::precalculate_matrices                  src/matrix_cache.cpp:121-171 (row 0 = e0, saturation)
//
// P(s -> c) = clamp01( sum_{j=0}^{min(s,c)} exp(t_j) * coeff^j ),
//   t_j = lnC(s, j) + lnC(s+c-1-j, s-1) + (s+c-2j) * ln(alpha)
// The reference evaluates t_j from a 1024-entry table of libm lgamma values (libm itself beyond it,
// src/probability.cpp:58-64); the host uploads those values for every argument up to 2 N, ln(alpha), coeff and the pow(coeff, j) row, so every input of exp() is bit-identical to
// the reference's and the terms are added in the same ascending-j order without FMA contraction.
// Only exp() itself (CUDA libdevice vs glibc, both < 1 ulp) can differ.
//
// Bound: FP64 ALU / transcendental (one exp per term, ~N^3/3 terms per matrix); not tensor, not HBM.
//
// Outputs, per unique key u (two layouts of the same numbers, columns c <= max_family_size only):
//   mp[u][c/4][s][c%4]   "panelised": a K-chunk of the matrix is one contiguous block, so the pruning
//                        kernel stages it with one 1-D bulk (TMA) copy and reads MMA A-fragments
//                        (8 rows x 4 cols = 32 consecutive doubles) without bank conflicts
//   mt[u][c][s]          transposed: column c contiguous over parent sizes, for leaf gathers and Pupko
// Rows s >= N and panel padding are never written (buffers are zeroed once at create).
#pragma once

#include "common.cuh"

namespace cafe {

struct KeyParams {
    double log_alpha;
    double coeff;
    int saturated;     // 1 - 2*alpha < 0 : only (0,0) = 1 is set           src/matrix_cache.cpp:152-153
    int computable;    // coeff > 0 && coeff != 1                           src/probability.cpp:154
};

struct MatrixBuildParams {
    int n;             // matrix size N = max(mrf, mf) + 1
    int mf;            // columns kept: c <= mf
    int nr;            // padded rows
    int n_keys;
    int key_first;     // this launch builds keys key_first .. key_first + gridDim.y - 1 (a slab, when the build is distributed)
    int lg_len;        // entries of the lgamma table (2 N + 2)
    const KeyParams* keys;      // [n_keys]
    const double* powc;         // [n_keys][n]  pow(coeff, j) from the host libm
    const double* lgamma_tab;   // [lg_len]     lgamma(i) from the host libm
    double* mp;
    double* mt;
    size_t mp_stride;
    size_t mt_stride;
};

constexpr int MB_THREADS = 256;

// One thread per matrix entry; blockIdx.y = key.  Entries are visited column-block-wise so a warp
// holds 32 consecutive parent sizes s for one child size c: the mt store is coalesced and the
// per-lane term counts min(s,c)+1 are similar.
__global__ void __launch_bounds__(MB_THREADS) bd_matrix_kernel(const MatrixBuildParams p)
{
    extern __shared__ double sm[];
    double* lg = sm;                       // [lg_len]
    double* pw = sm + p.lg_len;            // [n]
    const int key = p.key_first + blockIdx.y;
    const KeyParams kp = p.keys[key];
    for (int i = threadIdx.x; i < p.lg_len; i += MB_THREADS) lg[i] = p.lgamma_tab[i];
    for (int i = threadIdx.x; i < p.n; i += MB_THREADS) pw[i] = p.powc[(size_t)key * p.n + i];
    __syncthreads();

    const int cols = p.mf + 1;
    const int s_blocks = (p.n + 31) / 32;
    const int total = s_blocks * 32 * cols;
    double* mp = p.mp + (size_t)key * p.mp_stride;
    double* mt = p.mt + (size_t)key * p.mt_stride;
    for (int idx = blockIdx.x * MB_THREADS + threadIdx.x; idx < total; idx += gridDim.x * MB_THREADS) {
        const int lane_s = idx & 31;
        const int rest = idx >> 5;
        const int c = rest % cols;
        const int s = (rest / cols) * 32 + lane_s;
        if (s >= p.n) continue;
        double value = 0.0;
        if (s == 0) {
            value = (c == 0) ? 1.0 : 0.0;                       // src/matrix_cache.cpp:70-77
        }
        else if (!kp.saturated && kp.computable) {
            const int m = s < c ? s : c;
            const double lg_s1 = lg[s + 1];
            const double lg_s = lg[s];
            double total_p = 0.0;
            for (int j = 0; j <= m; ++j) {
                // chooseln(s, j): 0 when j == 0, else lg[s+1] - lg[j+1] - lg[s-j+1]      src/probability.cpp:79-88
                const double a = (j == 0) ? 0.0 : __dsub_rn(__dsub_rn(lg_s1, lg[j + 1]), lg[s - j + 1]);
                // chooseln(s+c-1-j, s-1): 0 when s == 1, else lg[s+c-j] - lg[s] - lg[c-j+1]
                const double b = (s == 1) ? 0.0 : __dsub_rn(__dsub_rn(lg[s + c - j], lg_s), lg[c - j + 1]);
                const double t = __dadd_rn(__dadd_rn(a, b), __dmul_rn((double)(s + c - 2 * j), kp.log_alpha));
                total_p = __dadd_rn(total_p, __dmul_rn(exp(t), pw[j]));
            }
            value = fmax(fmin(total_p, 1.0), 0.0);              // src/probability.cpp:144
        }
        mt[(size_t)c * p.nr + s] = value;
        mp[((size_t)(c >> 2) * p.nr + s) * 4 + (c & 3)] = value;
    }
}

}  // namespace cafe
