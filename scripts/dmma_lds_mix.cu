// How much shared-memory operand traffic does the FP64 tensor pipe tolerate?  Each warp runs the inner loop of the
// pruning kernel in isolation: per K panel, NLDS LDS.64 fragment loads (conflict-free, 256 B per warp each) feeding
// 10 DMMA.8x8x4 (5 row blocks x 2 family blocks), fragments double-buffered in registers, no barriers, no copies.
// Prints TFLOP/s for NLDS = 0 (operands stay in registers), 2, 7 (the kernel's 40x16 warp tile), 12 (a 40x8 tile
// would need 6 per 5 DMMAs) and for 4 x LDS.128 (the 7-load traffic in wider loads), at 1, 2 and 3 warps per
// sub-partition.   Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_lds_mix.bin.so dmma_lds_mix.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int ITERS = 2048;

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double lds64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void lds128(uint32_t addr, double& x, double& y)
{
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr));
}

// MODE: number of LDS.64 per panel (0, 2, 7, 12) or -4 for 4 x LDS.128
template <int MODE>
__global__ void mix_kernel(double* out)
{
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + (uint32_t)(warp * 2048 + lane * 8);     // 256 B per warp-load, conflict-free
    double acc[5][2][2];
    for (int i = 0; i < 5; ++i) acc[i][0][0] = acc[i][0][1] = acc[i][1][0] = acc[i][1][1] = 0.0;
    double a[2][5], b[2][2];
    for (int q = 0; q < 2; ++q) { for (int i = 0; i < 5; ++i) a[q][i] = 1.0 + lane * 1e-12 + i; b[q][0] = 0.5; b[q][1] = 0.25; }
    #pragma unroll 1
    for (int it = 0; it < ITERS; it += 2) {
        #pragma unroll
        for (int q = 0; q < 2; ++q) {           // panel q uses set q, loads refill set q^1 (double buffer)
            const uint32_t ad = base + (uint32_t)(((it + q) & 7) * 256 * 0);      // same lines every time: pure LSU/RF traffic, no capacity effects
            double* an = a[q ^ 1];
            double* bn = b[q ^ 1];
            if (MODE == 7 || MODE == 12) {
                #pragma unroll
                for (int i = 0; i < 5; ++i) an[i] = lds64(ad + i * 256);
                bn[0] = lds64(ad + 1280); bn[1] = lds64(ad + 1536);
                if (MODE == 12) {
                    #pragma unroll
                    for (int i = 0; i < 5; ++i) an[i] += 1e-30 * lds64(ad + 256 + i * 256);
                }
            }
            else if (MODE == 2) { bn[0] = lds64(ad); bn[1] = lds64(ad + 256); }
            else if (MODE == -4) {
                // 16-byte loads: the warp-load is 512 B; 4 of them carry the bytes of 8 LDS.64
                const uint32_t ad2 = base + lane * 8;      // lane * 16 in total
                lds128(ad2, an[0], an[1]); lds128(ad2 + 512, an[2], an[3]); lds128(ad2 + 1024, an[4], bn[0]);
                double t; lds128(ad2 + 1536, bn[1], t); bn[1] += 1e-30 * t;
            }
            #pragma unroll
            for (int i = 0; i < 5; ++i) {
                dmma(acc[i][0][0], acc[i][0][1], a[q][i], b[q][0]);
                dmma(acc[i][1][0], acc[i][1][1], a[q][i], b[q][1]);
            }
        }
    }
    double s = 0;
    for (int i = 0; i < 5; ++i) s += acc[i][0][0] + acc[i][0][1] + acc[i][1][0] + acc[i][1][1];
    if (s == 12345.678) out[0] = s;
}

template <int MODE>
double run(int sms, int warps, double* out)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const size_t smem = 4096 * sizeof(double);
    mix_kernel<MODE><<<sms, warps * 32, smem>>>(out);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0));
        mix_kernel<MODE><<<sms, warps * 32, smem>>>(out);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, (double)ms);
    }
    const double flops = 2.0 * 8 * 8 * 4 * 10 * ITERS * (double)sms * warps;
    return flops / best / 1e9;
}

int main()
{
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    double* out; CK(cudaMalloc(&out, 64));
    printf("{\"sms\": %d", sms);
    for (int warps : {4, 8, 12}) {
        printf(", \"w%d\": {\"lds0\": %.2f", warps, run<0>(sms, warps, out));
        printf(", \"lds2\": %.2f", run<2>(sms, warps, out));
        printf(", \"lds7\": %.2f", run<7>(sms, warps, out));
        printf(", \"lds12\": %.2f", run<12>(sms, warps, out));
        printf(", \"lds128x4\": %.2f}", run<-4>(sms, warps, out));
    }
    printf("}\n");
    return 0;
}
