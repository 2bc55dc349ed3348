// tmem_ld_microbench.cu -- what does a tcgen05.ld cost, alone and next to tcgen05.mma?
// One warp issues M128 x N64 x K16 MMAs back to back; 16 other warps (4 per TMEM lane quadrant)
// run tcgen05.ld loops over other TMEM columns.  Reported: cycles per load instruction with the
// loaders alone, cycles per MMA alone, and both when they run together -- i.e. whether accumulator
// flushes (conv_tc_acc* kernels) take tensor-pipe time from the MMAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tmem_ld_mb profiles/tmem_ld_microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}

// SHAPE: columns per load instruction (8, 16, 32)
template <int SHAPE>
__device__ __forceinline__ uint32_t ld_cols(uint32_t taddr) {
    uint32_t s = 0;
    if (SHAPE == 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]) :: "memory");
        for (int k = 0; k < 8; ++k) s ^= v[k];
    } else if (SHAPE == 16) {
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                       "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]) :: "memory");
        for (int k = 0; k < 16; ++k) s ^= v[k];
    } else {
        uint32_t v[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                       "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                       "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int k = 0; k < 32; ++k) s ^= v[k];
    }
    return s;
}

// mode bit 0: MMAs run, bit 1: loaders run
template <int SHAPE>
__global__ void bench(int mode, int mma_iters, int ld_iters, int ld_warps, long long* out, uint32_t* sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_s;
    __shared__ long long t_ld[16];
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
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
    long long t_mma = 0;
    if (warp == 16) {
        if (lane == 0 && (mode & 1)) {
            constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint64_t ad0 = desc(smem_u32(smem), 9792, 288), bd0 = desc(smem_u32(smem + 40 * 1024), 64 * 16, 128);
            const long long t0 = clock64();
            for (int it = 0; it < mma_iters; ++it) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    mma(tmem + (uint32_t)((j % 4) * 64), ad0 + (uint64_t)((j % 3) + 18 * (j % 5)), bd0 + (uint64_t)((j % 9) * 128), idesc, (it > 0 || j >= 4) ? 1u : 0u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            uint32_t done;
            do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
            } while (!done);
            t_mma = clock64() - t0;
        }
    } else if ((mode & 2) && warp < ld_warps) {
        const uint32_t taddr = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + 256u + (uint32_t)((warp >> 2) * 32);
        uint32_t s = 0;
        const long long t0 = clock64();
        for (int it = 0; it < ld_iters; ++it) s ^= ld_cols<SHAPE>(taddr + (uint32_t)((it & 1) * (SHAPE < 32 ? SHAPE : 0)));
        const long long t1 = clock64();
        if (lane == 0) t_ld[warp] = t1 - t0;
        if (s == 0xdeadbeefu) sink[0] = s;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (threadIdx.x == 0) {
        long long mx = 0;
        for (int w = 0; w < ld_warps; ++w) mx = t_ld[w] > mx ? t_ld[w] : mx;
        out[2 * blockIdx.x + 1] = (mode & 2) ? mx : 0;
    }
    if (warp == 16 && lane == 0) out[2 * blockIdx.x] = t_mma;
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int SHAPE>
void run(int mode, int mma_iters, int ld_iters, int ld_warps, long long* out, uint32_t* sink) {
    const int grid = 148;
    cudaFuncSetAttribute(bench<SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    bench<SHAPE><<<grid, 17 * 32, 64 * 1024>>>(mode, mma_iters, ld_iters, ld_warps, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("shape %d mode %d: %s\n", SHAPE, mode, cudaGetErrorString(e)); exit(1); }
    long long h[2 * 148];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    long long m0 = 0, m1 = 0;
    for (int i = 0; i < grid; ++i) { if (h[2 * i] > m0) m0 = h[2 * i]; if (h[2 * i + 1] > m1) m1 = h[2 * i + 1]; }
    printf("x%-2d  mma %s  loaders %2d warps x %5d loads : ", SHAPE, (mode & 1) ? "on " : "off", (mode & 2) ? ld_warps : 0, ld_iters);
    if (mode & 1) printf("%6.1f cycles/MMA (%d MMAs)   ", (double)m0 / (mma_iters * 16), mma_iters * 16);
    if (mode & 2) printf("%6.1f cycles per load per warp; SM total %5.2f cycles per load instr, %5.1f B/clk",
                         (double)m1 / ld_iters, (double)m1 / ((double)ld_iters * ld_warps), (double)ld_iters * ld_warps * SHAPE * 128.0 / m1);
    printf("\n");
}

int main() {
    long long* out; uint32_t* sink;
    cudaMalloc(&out, 2 * 148 * sizeof(long long)); cudaMalloc(&sink, 4);
    run<8>(1, 256, 0, 0, out, sink);
    for (int w : {1, 4, 16}) { run<8>(2, 0, 4096, w, out, sink); run<16>(2, 0, 4096, w, out, sink); run<32>(2, 0, 4096, w, out, sink); }
    // both: the MMAs take ~256*16*48 = 197k cycles alone; the loaders are sized to run about as long
    run<8>(3, 256, 1024, 16, out, sink); run<8>(3, 256, 2048, 16, out, sink); run<8>(3, 256, 4096, 16, out, sink);
    run<16>(3, 256, 1024, 16, out, sink); run<16>(3, 256, 2048, 16, out, sink);
    run<32>(3, 256, 512, 16, out, sink); run<32>(3, 256, 1024, 16, out, sink);
    run<8>(3, 256, 4096, 4, out, sink); run<32>(3, 256, 1024, 4, out, sink);
    return 0;
}
