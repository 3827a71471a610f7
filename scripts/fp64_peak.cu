// FP64 roofline denominators for the pruning kernel, measured on the box the bench runs on:
//   * DFMA issue peak (FP64 ALU)
//   * DMMA peak for every mma.sync f64 shape (m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16), register-resident
//   * cuBLAS DGEMM 8192^3 (the "library" FP64 GEMM peak, analogous to MEASURED_PEAKS.json's bf16 figure)
// MEASURED_PEAKS.json (driver-written) has HBM and bf16 only; the FP64 denominators come from here.
// Prints one JSON object.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 fp64_peak.cu -lcublas
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int ITERS = 4096;

__global__ void dfma_kernel(double* out, double a, double b)
{
    double x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;
}

template <int ACC>
__global__ void dmma_m8n8k4_kernel(double* out, double a, double b)
{
    double c[ACC][2];
    for (int i = 0; i < ACC; ++i) c[i][0] = c[i][1] = 0.0;
    double fa = a + threadIdx.x * 1e-12, fb = b;
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < ACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(fa), "d"(fb));
    }
    double s = 0;
    for (int i = 0; i < ACC; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

template <int ACC>
__global__ void dmma_m16n8k4_kernel(double* out, double a, double b)
{
    double c[ACC][4];
    for (int i = 0; i < ACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
    double fa0 = a + threadIdx.x * 1e-12, fa1 = a, fb = b;
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < ACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(fa0), "d"(fa1), "d"(fb));
    }
    double s = 0;
    for (int i = 0; i < ACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678) out[0] = s;
}

template <int ACC>
__global__ void dmma_m16n8k8_kernel(double* out, double a, double b)
{
    double c[ACC][4];
    for (int i = 0; i < ACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
    double fa0 = a + threadIdx.x * 1e-12, fa1 = a, fa2 = a * 0.5, fa3 = a * 0.25, fb0 = b, fb1 = b * 0.5;
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < ACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(fa0), "d"(fa1), "d"(fa2), "d"(fa3), "d"(fb0), "d"(fb1));
    }
    double s = 0;
    for (int i = 0; i < ACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678) out[0] = s;
}

template <int ACC>
__global__ void dmma_m16n8k16_kernel(double* out, double a, double b)
{
    double c[ACC][4];
    for (int i = 0; i < ACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
    double fa[8], fb[4];
    for (int i = 0; i < 8; ++i) fa[i] = a + threadIdx.x * 1e-12 + i * 1e-3;
    for (int i = 0; i < 4; ++i) fb[i] = b + i * 1e-3;
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < ACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(fa[0]), "d"(fa[1]), "d"(fa[2]), "d"(fa[3]), "d"(fa[4]), "d"(fa[5]), "d"(fa[6]), "d"(fa[7]),
                           "d"(fb[0]), "d"(fb[1]), "d"(fb[2]), "d"(fb[3]));
    }
    double s = 0;
    for (int i = 0; i < ACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678) out[0] = s;
}

template <typename F>
double time_ms(F launch, int reps = 5)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, (double)ms);
    }
    return best;
}

int main()
{
    int dev = 0, sms = 0, clk = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev));
    double* out; CK(cudaMalloc(&out, 64));
    printf("{\"sms\": %d, \"clock_khz\": %d", sms, clk);

    for (int warps : {4, 8, 16, 32}) {
        const int blocks = sms * 2, threads = warps * 32 / 2;
        double ms = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
        double flops = 2.0 * 8 * ITERS * (double)blocks * threads;
        printf(", \"dfma_tflops_w%d\": %.3f", warps, flops / ms / 1e9);
    }
    for (int warps : {4, 8, 16}) {
        const int blocks = sms, threads = warps * 32;
        double ms = time_ms([&] { dmma_m8n8k4_kernel<8><<<blocks, threads>>>(out, 1.0000001, 1e-9); });
        double flops = 2.0 * 8 * 8 * 4 * 8 * ITERS * (double)blocks * warps;
        printf(", \"dmma_m8n8k4_tflops_w%d\": %.3f", warps, flops / ms / 1e9);
        ms = time_ms([&] { dmma_m16n8k4_kernel<4><<<blocks, threads>>>(out, 1.0000001, 1e-9); });
        flops = 2.0 * 16 * 8 * 4 * 4 * ITERS * (double)blocks * warps;
        printf(", \"dmma_m16n8k4_tflops_w%d\": %.3f", warps, flops / ms / 1e9);
        ms = time_ms([&] { dmma_m16n8k8_kernel<4><<<blocks, threads>>>(out, 1.0000001, 1e-9); });
        flops = 2.0 * 16 * 8 * 8 * 4 * ITERS * (double)blocks * warps;
        printf(", \"dmma_m16n8k8_tflops_w%d\": %.3f", warps, flops / ms / 1e9);
        ms = time_ms([&] { dmma_m16n8k16_kernel<4><<<blocks, threads>>>(out, 1.0000001, 1e-9); });
        flops = 2.0 * 16 * 8 * 16 * 4 * ITERS * (double)blocks * warps;
        printf(", \"dmma_m16n8k16_tflops_w%d\": %.3f", warps, flops / ms / 1e9);
    }
    {
        // 2 accumulators only: dependent-issue latency of DMMA
        double ms = time_ms([&] { dmma_m8n8k4_kernel<1><<<sms, 32>>>(out, 1.0000001, 1e-9); });
        printf(", \"dmma_m8n8k4_latency_ns\": %.2f", ms * 1e6 / ITERS);
    }
    for (int n : {4096, 8192}) {
        cublasHandle_t h; cublasCreate(&h);
        double *A, *B, *C;
        size_t bytes = (size_t)n * n * sizeof(double);
        CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
        std::vector<double> hA((size_t)n * n);
        for (size_t i = 0; i < hA.size(); ++i) hA[i] = (double)((i * 2654435761u) % 1000) / 1000.0 - 0.5;
        CK(cudaMemcpy(A, hA.data(), bytes, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(B, hA.data(), bytes, cudaMemcpyHostToDevice));
        const double alpha = 1.0, beta = 0.0;
        double ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &alpha, A, n, B, n, &beta, C, n); }, 4);
        printf(", \"cublas_dgemm_%d_tflops\": %.3f", n, 2.0 * n * n * (double)n / ms / 1e9);
        if (n == 8192) {
            // sustained: back to back for ~3 s
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            int reps = std::max(4, (int)(3000.0 / ms));
            CK(cudaEventRecord(e0));
            for (int r = 0; r < reps; ++r) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &alpha, A, n, B, n, &beta, C, n);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float tot; CK(cudaEventElapsedTime(&tot, e0, e1));
            printf(", \"cublas_dgemm_8192_sustained_tflops\": %.3f", 2.0 * n * n * (double)n * reps / tot / 1e9);
        }
        cudaFree(A); cudaFree(B); cudaFree(C); cublasDestroy(h);
    }
    printf("}\n");
    return 0;
}
