// Kernel 3b — per-family finalisation and the deterministic reduction of sum_i lnL_i.
//
// Restates (file:line in the reference)
//   gamma: family_likelihood = accumulate(cat_likelihoods), lnL = log(.)        src/gamma_core.cpp:207,218
//          any failed family -> the evaluation returns -log(0)                    src/gamma_core.cpp:227-236
//   base : lnL_i already produced by the root op; score = -accumulate(lnL)        src/base_model.cpp:107
//
// The reference adds the per-family values serially in family order; here each block adds its
// slice in a fixed tree order and one block adds the block partials in index order, so the result
// is run-to-run deterministic and differs from the serial sum only by reassociation (~1e-16 rel).
// Bound: HBM (reads k doubles + k flags, writes one double per family) — a few MB, latency-sized.
#pragma once

#include <climits>

#include "common.cuh"

namespace cafe {

constexpr int RED_THREADS = 256;

// cat_lk / fail arrive category-major [k][F] (what a category pass of the pruning kernel writes contiguously); the
// family-major [F][k] copy the callers read (gamma_model::_category_likelihoods) is written here.
__global__ void __launch_bounds__(RED_THREADS) finalize_kernel(int64_t n_families, int k, int mode, const double* __restrict__ cat_lk,
                                                               const uint8_t* __restrict__ fail, double* __restrict__ family_lnl,
                                                               uint8_t* __restrict__ family_fail, double* __restrict__ partial,
                                                               double* __restrict__ cat_lk_fm)
{
    __shared__ double s_sum[RED_THREADS];
    __shared__ double s_bad[RED_THREADS];
    double sum = 0.0, bad = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * RED_THREADS + threadIdx.x; i < n_families; i += (int64_t)gridDim.x * RED_THREADS) {
        double lnl;
        bool failed = false;
        if (mode == 0) lnl = cat_lk[i];
        else {
            double fam = 0.0;
            for (int c = 0; c < k; ++c) {
                failed |= fail[(size_t)c * n_families + i] != 0;
                const double v = cat_lk[(size_t)c * n_families + i];
                fam += v;                                     // ascending category, as std::accumulate (src/gamma_core.cpp:207)
                cat_lk_fm[i * k + c] = v;
            }
            lnl = log(fam);
        }
        if (failed) { lnl = __longlong_as_double(0x7ff8000000000000LL); bad += 1.0; }
        else sum += lnl;
        family_lnl[i] = lnl;
        family_fail[i] = failed ? 1 : 0;
    }
    s_sum[threadIdx.x] = sum;
    s_bad[threadIdx.x] = bad;
    __syncthreads();
    for (int off = RED_THREADS / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            s_sum[threadIdx.x] += s_sum[threadIdx.x + off];
            s_bad[threadIdx.x] += s_bad[threadIdx.x + off];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = s_sum[0];
        partial[2 * blockIdx.x + 1] = s_bad[0];
    }
}

__global__ void __launch_bounds__(RED_THREADS) final_sum_kernel(int n_partials, const double* __restrict__ partial, double* __restrict__ result)
{
    __shared__ double s_sum[RED_THREADS];
    __shared__ double s_bad[RED_THREADS];
    double sum = 0.0, bad = 0.0;
    for (int i = threadIdx.x; i < n_partials; i += RED_THREADS) { sum += partial[2 * i]; bad += partial[2 * i + 1]; }
    s_sum[threadIdx.x] = sum;
    s_bad[threadIdx.x] = bad;
    __syncthreads();
    for (int off = RED_THREADS / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            s_sum[threadIdx.x] += s_sum[threadIdx.x + off];
            s_bad[threadIdx.x] += s_bad[threadIdx.x + off];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { result[0] = s_sum[0]; result[1] = s_bad[0]; }
}

// Ingest of an uploaded count matrix (cafe_b200_set_families): narrow the caller's elements (1, 2 or 4 bytes) to the
// device width (1 byte when max_family_size <= 255, else 2) and track range[0] = min, range[1] = max over all leaf
// counts in the same pass, so that the reference's out-of-range indexing (src/probability.cpp:191,197) is refused
// without a host pass over the matrix.  Out-of-range values are stored as 0 (the context refuses to evaluate them).
// Bound: HBM — one read (and one narrower write) per count, 16-byte loads on the common paths, grid-stride.
__global__ void __launch_bounds__(RED_THREADS) ingest_counts_kernel(const void* src, int src_bytes, void* dst, int dst_bytes,
                                                                    int64_t n, int mf, int* __restrict__ range)
{
    int lo = INT_MAX, hi = INT_MIN;
    const int64_t tid0 = (int64_t)blockIdx.x * RED_THREADS + threadIdx.x, stride = (int64_t)gridDim.x * RED_THREADS;
    int64_t done = 0;                    // elements handled by the 16-byte paths below
    if (src_bytes == 1 && dst_bytes == 1 && src == dst) {
        // one byte per count already in place (the common upload): range check only, 16 counts per load
        const uint4* v = reinterpret_cast<const uint4*>(src);
        const int64_t n16 = n / 16;
        unsigned lo4 = 0xffffffffu, hi4 = 0u;
        for (int64_t i = tid0; i < n16; i += stride) {
            const uint4 x = __ldg(v + i);
            lo4 = __vminu4(__vminu4(lo4, x.x), __vminu4(__vminu4(x.y, x.z), x.w));
            hi4 = __vmaxu4(__vmaxu4(hi4, x.x), __vmaxu4(__vmaxu4(x.y, x.z), x.w));
        }
        if (tid0 < n16) {
            lo = min(min((int)(lo4 & 0xff), (int)((lo4 >> 8) & 0xff)), min((int)((lo4 >> 16) & 0xff), (int)(lo4 >> 24)));
            hi = max(max((int)(hi4 & 0xff), (int)((hi4 >> 8) & 0xff)), max((int)((hi4 >> 16) & 0xff), (int)(hi4 >> 24)));
        }
        done = n16 * 16;
    }
    else if (src_bytes == 4 && dst_bytes == 1) {
        // int32 counts narrowed to one byte: four counts per 16-byte load and 4-byte store
        const int4* v = reinterpret_cast<const int4*>(src);
        uchar4* out = reinterpret_cast<uchar4*>(dst);
        const int64_t n4 = n / 4;
        for (int64_t i = tid0; i < n4; i += stride) {
            const int4 x = __ldg(v + i);
            lo = min(lo, min(min(x.x, x.y), min(x.z, x.w)));
            hi = max(hi, max(max(x.x, x.y), max(x.z, x.w)));
            auto nar = [mf](int c) { return (unsigned char)((c < 0 || c > mf) ? 0 : c); };
            out[i] = make_uchar4(nar(x.x), nar(x.y), nar(x.z), nar(x.w));
        }
        done = n4 * 4;
    }
    for (int64_t i = done + tid0; i < n; i += stride) {
        int v;
        if (src_bytes == 4) v = reinterpret_cast<const int32_t*>(src)[i];
        else if (src_bytes == 2) v = reinterpret_cast<const uint16_t*>(src)[i];
        else v = reinterpret_cast<const uint8_t*>(src)[i];
        lo = min(lo, v);
        hi = max(hi, v);
        if (src != dst || src_bytes != dst_bytes) {
            const int w = (v < 0 || v > mf) ? 0 : v;
            if (dst_bytes == 1) reinterpret_cast<uint8_t*>(dst)[i] = (uint8_t)w;
            else reinterpret_cast<uint16_t*>(dst)[i] = (uint16_t)w;
        }
    }
    #pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, off));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, off));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&range[0], lo);
        atomicMax(&range[1], hi);
    }
}

}  // namespace cafe
