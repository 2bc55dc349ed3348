// umma_rounding_test.cu -- how does tcgen05.mma (kind::f16, fp32 accumulate) round when it adds a
// k-step's product sum into the TMEM accumulator?  D starts at 1.0 (first MMA: 1*1), then S MMAs
// each add exactly 0.75 ulp(1.0) = 1.5 * 2^-24 (= (1.5*2^-12) * 2^-12, exact in fp32).
//   round-to-nearest : every add rounds up by one ulp      -> 1 + S ulp
//   round-toward-zero: every add is truncated away         -> 1.0
//   exact            :                                        1 + 0.75 S ulp
// Second experiment: 16 such products inside ONE k-step (sum = 12 ulp exactly) to see the in-step adder.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// A images: [k-chunk(2)][row(128)][8]; B images: [k-chunk(2)][n(16)][8]
__global__ void test(int S, float* out) {
    __shared__ __align__(1024) __half a_one[2 * 128 * 8], a_small[2 * 128 * 8], a_small16[2 * 128 * 8];
    __shared__ __align__(128) __half b_one[2 * 16 * 8], b_small[2 * 16 * 8], b_small16[2 * 16 * 8];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_s;
    for (int i = threadIdx.x; i < 2 * 128 * 8; i += blockDim.x) {
        const int k = (i / (128 * 8)) * 8 + (i % 8);      // k index 0..15
        a_one[i] = __float2half(k == 0 ? 1.f : 0.f);
        a_small[i] = __float2half(k == 0 ? 1.5f * 0.000244140625f : 0.f);   // 1.5 * 2^-12
        a_small16[i] = __float2half(1.5f * 0.000244140625f);
    }
    for (int i = threadIdx.x; i < 2 * 16 * 8; i += blockDim.x) {
        const int k = (i / (16 * 8)) * 8 + (i % 8);
        b_one[i] = __float2half(k == 0 ? 1.f : 0.f);
        b_small[i] = __float2half(k == 0 ? 0.000244140625f : 0.f);          // 2^-12
        b_small16[i] = __float2half(0.000244140625f);
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(32));
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
    const uint32_t idesc = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (threadIdx.x == 0) {
        auto mma = [&](const __half* a, const __half* b, uint32_t col, uint32_t acc) {
            const uint64_t ad = desc(smem_u32(a), 128 * 16, 128), bd = desc(smem_u32(b), 16 * 16, 128);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem + col), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        };
        mma(a_one, b_one, 0, 0);
        for (int i = 0; i < S; ++i) mma(a_small, b_small, 0, 1);
        mma(a_one, b_one, 16, 0);
        mma(a_small16, b_small16, 16, 1);       // one k-step adding 16 * 0.75 ulp = 12 ulp
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    } while (!done);
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (threadIdx.x < 32) {
        uint32_t v0, v1;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v0) : "r"(tmem));
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v1) : "r"(tmem + 16));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (threadIdx.x == 0) { out[0] = __uint_as_float(v0); out[1] = __uint_as_float(v1); }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
}

int main() {
    float* out; cudaMalloc(&out, 8);
    for (int S : {1, 2, 4, 16, 64}) {
        test<<<1, 128>>>(S, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        float h[2]; cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
        const double ulp = ldexp(1.0, -23);
        printf("S=%3d chained adds of 0.75 ulp: D = 1 + %.2f ulp   (RN: %d ulp, RZ: 0 ulp, exact: %.2f ulp)   |  one k-step of 16 x 0.75 ulp: D = 1 + %.2f ulp (exact 12)\n",
               S, (h[0] - 1.0) / ulp, S, 0.75 * S, (h[1] - 1.0) / ulp);
    }
    return 0;
}
