// Kernel 5 — family p-values from simulated conditional distributions.
//
// Restates (file:line in the reference)
//   get_random_probabilities: sort of the simulated likelihoods   src/probability.cpp:310
//   pvalue (upper_bound index / size)                             src/probability.cpp:379-389
//   compute_tree_pvalue (max over root sizes)                     src/probability.cpp:391-409
//
// The likelihoods themselves (max over the root vector, simulated and observed families alike) come from
// the pruning kernel in mode 2.  Bound: HBM / latency — n_root_sizes x n_sim doubles sorted once per lambda,
// then n_families x n_root_sizes binary searches over L2-resident rows.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace cafe {

constexpr int PV_THREADS = 256;
constexpr int PV_MAX_SIM = 4096;        // one row is sorted in shared memory (32 KB)

// One block per root size: bitonic sort of its n_sim values (ascending), padded with +inf to a power of two.
__global__ void __launch_bounds__(PV_THREADS) sort_rows_kernel(double* cond, int n_sim, int padded)
{
    extern __shared__ double row_s[];
    double* row = cond + (size_t)blockIdx.x * n_sim;
    for (int i = threadIdx.x; i < padded; i += PV_THREADS) row_s[i] = i < n_sim ? row[i] : INFINITY;
    __syncthreads();
    for (int k = 2; k <= padded; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < padded; i += PV_THREADS) {
                const int partner = i ^ j;
                if (partner > i) {
                    const double x = row_s[i], y = row_s[partner];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { row_s[i] = y; row_s[partner] = x; }
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < n_sim; i += PV_THREADS) row[i] = row_s[i];
}

// One thread per family: p = max_s idx_s / n_sim with idx_s = number of simulated values <= observed, or
// n_sim - 1 when none is greater (src/probability.cpp:381-388).
__global__ void __launch_bounds__(PV_THREADS) pvalue_kernel(const double* __restrict__ cond, int n_root_sizes, int n_sim,
                                                            const double* __restrict__ observed, int64_t n_families, double* __restrict__ pvalues)
{
    const int64_t f = (int64_t)blockIdx.x * PV_THREADS + threadIdx.x;
    if (f >= n_families) return;
    const double v = observed[f];
    int best = 0;
    for (int s = 0; s < n_root_sizes; ++s) {
        const double* row = cond + (size_t)s * n_sim;
        int lo = 0, hi = n_sim;             // first index with row[idx] > v
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(row + mid) > v) hi = mid;
            else lo = mid + 1;
        }
        const int idx = lo < n_sim ? lo : n_sim - 1;
        best = max(best, idx);
    }
    pvalues[f] = best / (double)n_sim;      // max of idx/size over s == (max idx)/size: the division is monotone
}

}  // namespace cafe
