// umma_microbench.cu -- cycles per tcgen05.mma (kind::f16, M=128, cta_group::1) as a function
// of N, of the shared-memory operand layout (no swizzle vs 128B swizzle) and of how many
// independent TMEM accumulators the issuing thread round-robins over.  The issue loop is
// fully unrolled with precomputed descriptors so that the single issuing thread is not the
// bottleneck (a first version with per-MMA descriptor arithmetic measured ~110 cycles per
// MMA for every N: that was the scalar issue loop, not the tensor pipe).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_mb profiles/umma_microbench.cu && /tmp/umma_mb
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}

// MODE 0: no swizzle (A: SBO=288 pool layout, LBO=9792; B canonical); MODE 1: no swizzle, dense A
// (SBO=128, LBO=2048: the 128 rows of a k-chunk contiguous, as final_tapsum_kernel's); MODE 2: 128B swizzle
template <int N, int MODE, int NT>
__global__ void bench(int iters, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_s;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_s;
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 48 * 1024);
    if (threadIdx.x == 0) {
        const uint64_t ad0 = MODE == 0 ? desc(a0, 9792, 288, 0) : MODE == 1 ? desc(a0, 2048, 128, 0) : desc(a0, 16, 1024, 2);
        const uint64_t bd0 = MODE != 2 ? desc(b0, N * 16, 128, 0) : desc(b0, 16, 1024, 2);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                // constant per-MMA offsets (16-byte units) folded at compile time
                const uint64_t ad = ad0 + (uint64_t)(MODE == 0 ? (j % 3) + 18 * (j % 5) : MODE == 1 ? 256 * (j % 8) : 2 * (j % 4));
                const uint64_t bd = bd0 + (uint64_t)(MODE != 2 ? (j % (N >= 256 ? 3 : 9)) * N * 2 : 2 * (j % 4));
                mma(tmem + (uint32_t)((j % NT) * N), ad, bd, idesc, (it > 0 || j >= NT) ? 1u : 0u);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        } while (!done);
        out[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int N, int MODE, int NT>
void run(long long* out) {
    const int iters = 128, grid = 148;
    cudaFuncSetAttribute(bench<N, MODE, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    bench<N, MODE, NT><<<grid, 128, 100 * 1024>>>(iters, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N %d mode %d: %s\n", N, MODE, cudaGetErrorString(e)); exit(1); }
    long long h[148];
    cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%s  N=%3d  accumulators=%d : %6.1f cycles/MMA  (tensor ideal %5.1f; smem A+B read %5.1f)\n",
           MODE == 0 ? "no-swizzle pool " : MODE == 1 ? "no-swizzle dense" : "swizzle128      ", N, NT, (double)mx / (iters * 16), N / 2.0, (4096 + N * 32) / 128.0);
}

int main() {
    long long* out;
    cudaMalloc(&out, 148 * sizeof(long long));
    run<16, 0, 1>(out); run<16, 0, 4>(out);
    run<32, 0, 1>(out); run<32, 0, 2>(out); run<32, 0, 4>(out); run<32, 0, 8>(out);
    run<64, 0, 1>(out); run<64, 0, 2>(out); run<64, 0, 4>(out); run<64, 0, 8>(out);
    run<128, 0, 1>(out); run<128, 0, 2>(out); run<128, 0, 4>(out);
    run<256, 0, 1>(out); run<256, 0, 2>(out);
    run<16, 1, 4>(out); run<32, 1, 4>(out); run<64, 1, 4>(out); run<128, 1, 4>(out); run<256, 1, 2>(out);
    run<32, 2, 1>(out); run<32, 2, 4>(out); run<64, 2, 1>(out); run<64, 2, 4>(out);
    run<128, 2, 1>(out); run<128, 2, 4>(out); run<256, 2, 1>(out); run<256, 2, 2>(out);
    return 0;
}
