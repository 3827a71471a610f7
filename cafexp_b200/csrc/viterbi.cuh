// Kernel 6 — branch probabilities of the reconstructed size changes ("Viterbi sums").
//
// Restates (file:line in the reference)
//   compute_viterbi_sum      src/gene_family_reconstructor.cpp:361-400   (caller src/execute.cpp:165-176)
//
// For a family and a non-root node whose reconstructed size differs from its parent's: with p* = M[parent][child],
// the sum over m = 0 .. max_family_size-1, ascending, of  M[parent][m]/2 where M[parent][m] == p*  and  M[parent][m]
// where M[parent][m] < p*.  Same size as the parent, or the root: no value (reported as -1).
// One thread per (family, node); the row scan is sequential per thread so the sum has the reference's order.
// The matrices are the device-resident ones of the last build (transposed layout: column m contiguous over rows).
// Bound: L2 latency — (selected families) x (nodes) x max_family_size reads of 8 B, a few MB.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace cafe {

constexpr int VT_THREADS = 128;

__global__ void __launch_bounds__(VT_THREADS) viterbi_kernel(int64_t n_families, int n_nodes, int mf, int nr, const int* __restrict__ parent,
                                                             const int* __restrict__ mat_of, const double* __restrict__ mt, size_t mt_stride,
                                                             const int32_t* __restrict__ node_sizes, const uint8_t* __restrict__ selected,
                                                             double* __restrict__ out)
{
    const int64_t idx = (int64_t)blockIdx.x * VT_THREADS + threadIdx.x;
    if (idx >= n_families * n_nodes) return;
    const int64_t f = idx / n_nodes;
    const int v = (int)(idx - f * n_nodes);
    double result = -1.0;
    const int par = parent[v];
    if (par >= 0 && (selected == nullptr || selected[f])) {
        const int ps = node_sizes[f * n_nodes + par], cs = node_sizes[f * n_nodes + v];
        if (ps != cs) {
            const double* m = mt + (size_t)mat_of[v] * mt_stride + ps;      // M[ps][c] = m[c * nr]
            const double pstar = __ldg(m + (size_t)cs * nr);
            double acc = 0.0;
            for (int c = 0; c < mf; ++c) {
                const double pm = __ldg(m + (size_t)c * nr);
                if (pm == pstar) acc = __dadd_rn(acc, pm / 2.0);
                else if (pm < pstar) acc = __dadd_rn(acc, pm);
            }
            result = acc;
        }
    }
    out[idx] = result;
}

}  // namespace cafe
