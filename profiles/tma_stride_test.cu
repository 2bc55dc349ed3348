// tma_stride_test.cu -- pins down the semantics of cuTensorMapEncodeTiled elementStrides on
// sm_100a before the L2 kernel relies on them: a rank-4 fp16 tensor [c][y][x][8] is read
// with traversal stride 2 along x starting at x = -1 (out of bounds -> zero fill) and
// y = -1, and the units that landed in shared memory are printed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_stride_test tma_stride_test.cu
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int C = 2, Y = 6, X = 8;
constexpr int SM_UNITS = 1024;

__global__ void probe(const __grid_constant__ CUtensorMap tm, int x0, int y0, uint32_t expect_bytes, uint4* out,
                      int* status) {
    __shared__ __align__(128) uint4 buf[SM_UNITS];
    __shared__ __align__(8) uint64_t bar;
    for (int i = threadIdx.x; i < SM_UNITS; i += blockDim.x) buf[i] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(expect_bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2,%3,%4,%5}], [%6];"
            ::"r"((uint32_t)__cvta_generic_to_shared(buf)), "l"(&tm), "r"(0), "r"(x0), "r"(y0), "r"(0), "r"(b) : "memory");
        uint32_t done = 0;
        const long long t0 = clock64();
        while (!done && clock64() - t0 < 20000000LL) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0,1,0,p;\n\t}"
                         : "=r"(done) : "r"(b) : "memory");
        }
        *status = (int)done;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SM_UNITS; i += blockDim.x) out[i] = buf[i];
}

int main() {
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres) != cudaSuccess || !encode) {
        printf("no cuTensorMapEncodeTiled\n");
        return 1;
    }
    std::vector<__half> h((size_t)C * Y * X * 8);
    for (int c = 0; c < C; ++c)
        for (int y = 0; y < Y; ++y)
            for (int x = 0; x < X; ++x)
                for (int k = 0; k < 8; ++k) h[(((size_t)c * Y + y) * X + x) * 8 + k] = __float2half((float)(c * 1000 + y * 10 + x));
    __half* d;
    cudaMalloc(&d, h.size() * 2);
    cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    uint4* out;
    int* status;
    cudaMalloc(&out, SM_UNITS * 16);
    cudaMalloc(&status, 4);

    for (int boxx : {9, 10}) {
        CUtensorMap tm;
        cuuint64_t dims[4] = {8, X, Y, C};
        cuuint64_t strides[3] = {16, 16 * X, 16 * X * Y};
        cuuint32_t box[4] = {8, (cuuint32_t)boxx, (cuuint32_t)(Y + 2), C};
        cuuint32_t estr[4] = {1, 2, 1, 1};
        CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("boxDim[x]=%d elementStride 2: encode rc=%d\n", boxx, (int)r);
        if (r != CUDA_SUCCESS) continue;
        for (int nx : {4, 5}) {
            const uint32_t expect = (uint32_t)(nx * (Y + 2) * C * 16);
            cudaMemset(status, 0, 4);
            probe<<<1, 128>>>(tm, -1, -1, expect, out, status);
            cudaError_t e = cudaDeviceSynchronize();
            int st = 0;
            cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost);
            std::vector<uint4> o(SM_UNITS);
            cudaMemcpy(o.data(), out, SM_UNITS * 16, cudaMemcpyDeviceToHost);
            int landed = 0;
            for (int i = 0; i < SM_UNITS; ++i) landed += o[i].x != 0xFFFFFFFFu;
            printf("  expect %d units/row (%u B): barrier %s, %d units landed, err=%s\n", nx, expect,
                   st ? "completed" : "TIMED OUT", landed, cudaGetErrorString(e));
            if (st) {
                const int per_row = landed / ((Y + 2) * C);
                for (int row = 0; row < 3; ++row) {
                    printf("    smem row %d:", row);
                    for (int u = 0; u < per_row; ++u) {
                        const __half* hv = reinterpret_cast<const __half*>(&o[row * per_row + u]);
                        printf(" %g", __half2float(hv[0]));
                    }
                    printf("\n");
                }
                const int r2 = (Y + 2) + 1;   // c = 1, y = 0
                printf("    smem row %d:", r2);
                for (int u = 0; u < per_row; ++u) {
                    const __half* hv = reinterpret_cast<const __half*>(&o[r2 * per_row + u]);
                    printf(" %g", __half2float(hv[0]));
                }
                printf("\n");
                break;
            }
        }
    }
    return 0;
}
