// umma_layout_microbench.cu -- cycles per tcgen05.mma for the EXACT operand addressing of the layer-2
// kernels (conv_tc_acc2*): A = TMA-staged half block, planes [parity][chunk][34 rows][9 units of 16 B],
// LBO = one plane, SBO = two staged rows, start address shifted per filter tap and pooling phase;
// B = weight image.  Variants: row pitch (units per staged row), plane padding, N.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/umma_layout_mb profiles/umma_layout_microbench.cu
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

// ROW_UNITS: 16-byte units per staged row (9 in the kernel); PLANE_PAD: extra bytes per plane
template <int N, int ROW_UNITS, int PLANE_PAD, int NWARPS>
__global__ void bench(int iters, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_s;
    constexpr int ROW_B = ROW_UNITS * 16, PLANE_B = 34 * ROW_B + PLANE_PAD, PAR_B = 4 * PLANE_B, REGION_B = 2 * PAR_B;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(NWARPS));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_s;
    const long long t_start = clock64();
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (warp < NWARPS && lane == 0) {
        const int t = warp, py = t >> 1, px = t & 1;
        const uint32_t a_base = smem_u32(smem) + py * ROW_B;
        const uint64_t a_hi0 = desc(a_base, PLANE_B, 2 * ROW_B), a_lo0 = desc(a_base + REGION_B, PLANE_B, 2 * ROW_B);
        const uint64_t b0 = desc(smem_u32(smem) + 2 * REGION_B, N * 16, 128);
        uint64_t dxo[3];
        for (int dx = 0; dx < 3; ++dx) dxo[dx] = (uint64_t)((((px + dx) & 1) * PAR_B + ((px + dx) >> 1) * 16) >> 4);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap)
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const int dy = tap / 3, dx = tap % 3;
                    const uint64_t ad = ((tap & 1) ? a_lo0 : a_hi0) + dxo[dx] + (uint64_t)((dy * ROW_B + 2 * s * PLANE_B) >> 4);
                    const uint64_t bd = b0 + (uint64_t)(((tap * 4 + 2 * s) * N * 16) >> 4);
                    mma(tmem + (uint32_t)(t * (512 / NWARPS >= N ? N : 0)), ad, bd, idesc, (it > 0 || tap > 0 || s > 0) ? 1u : 0u);
                }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    if (threadIdx.x == 0) {
        const long long t0 = t_start;
        uint32_t done;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        } while (!done);
        out[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int N, int ROW_UNITS, int PLANE_PAD, int NWARPS>
void run(long long* out) {
    const int iters = 64, grid = 148;
    cudaFuncSetAttribute(bench<N, ROW_UNITS, PLANE_PAD, NWARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    bench<N, ROW_UNITS, PLANE_PAD, NWARPS><<<grid, 128, 200 * 1024>>>(iters, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N %d: %s\n", N, cudaGetErrorString(e)); exit(1); }
    long long h[148];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("N=%3d  row %2d units (%3d B)  plane pad %3d B (LBO %5d B, mod 128 = %3d)  %d issuing warps : %6.1f cycles/MMA\n", N, ROW_UNITS,
           ROW_UNITS * 16, PLANE_PAD, 34 * ROW_UNITS * 16 + PLANE_PAD, (34 * ROW_UNITS * 16 + PLANE_PAD) % 128, NWARPS,
           (double)mx / (iters * 18.0 * NWARPS));
}

int main() {
    long long* out;
    cudaMalloc(&out, 148 * sizeof(long long));
    run<64, 9, 0, 1>(out); run<64, 9, 0, 4>(out);          // the kernel's layout
    run<64, 9, 32, 4>(out); run<64, 9, 96, 4>(out); run<64, 9, 64, 4>(out);
    run<64, 10, 0, 4>(out); run<64, 12, 0, 4>(out); run<64, 16, 0, 4>(out); run<64, 8, 0, 4>(out);
    run<128, 9, 0, 2>(out); run<128, 9, 96, 2>(out); run<128, 8, 0, 2>(out);
    run<32, 9, 0, 4>(out);
    return 0;
}
