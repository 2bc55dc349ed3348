// fp64_peak_test.cu -- measured fp64 throughput of this B200: DFMA (CUDA cores) and DMMA
// (mma.sync m8n8k4 f64), the two candidate engines of the PCA projection (score.cu).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak_test fp64_peak_test.cu
#include <cuda_runtime.h>
#include <cstdio>

__global__ void dfma_kernel(double* out, int iters, double x) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], x, 1e-9);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 12345.678) out[0] = s;
}

__global__ void dmma_kernel(double* out, int iters, double x) {
    double d[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i][0] = d[i][1] = threadIdx.x * 1e-3 + i;
    const double a = x, b = x * 0.5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(d[i][0]), "+d"(d[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1];
    if (s == 12345.678) out[0] = s;
}

int main() {
    double* out;
    cudaMalloc(&out, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096, blocks = 148 * 8, threads = 256;
    for (int which = 0; which < 2; ++which) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (which == 0) dfma_kernel<<<blocks, threads>>>(out, iters, 0.999);
            else dmma_kernel<<<blocks, threads>>>(out, iters, 0.999);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double fma_per_thread = which == 0 ? 8.0 * iters : 8.0 * iters * 256.0 / 32.0;
            const double flops = 2.0 * fma_per_thread * blocks * threads;
            printf("%s: %.3f ms, %.2f TFLOP/s fp64\n", which == 0 ? "DFMA" : "DMMA m8n8k4", ms, flops / ms * 1e-9);
        }
    }
    return 0;
}
