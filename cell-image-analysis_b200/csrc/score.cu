// score.cu -- K4 RobustScaler -> PCA projection, K5 RBF one-class SVM decision for
// both detectors (exact fp64 direct-difference path), K6 per-strain accumulators.
//
// Replaces improved_detection.py:134-135 (scaler.transform, pca.transform), :138-142
// (predict + decision_function of the two OneClassSVMs; libsvm k_function RBF,
// sklearn/svm/src/libsvm/svm.cpp:461-472, decision sum - rho, sign rule sum > 0) and
// the reductions behind :151-152 and :202-211.
#include "common.cuh"

namespace {

constexpr int PT = 256;        // threads
constexpr int PB = 64;         // cells per block
constexpr int PFT = 32;        // features per shared-memory tile
constexpr int PCT = 128;       // components per block tile
constexpr int XPAD = PB + 2;   // padded row of the transposed x tile

// z[c] = sum_f x2[f] * Wt[f][c] - offset[c]: a register-blocked fp64 GEMM tile (64 cells x 128
// components per block, 8 x 4 accumulators per thread) over fp32-rounded inputs.
// x2 mirrors sklearn's in-place float32 flow: x1 = f32(x - center), x2 = f32(f64(x1)/scale).
__global__ void __launch_bounds__(PT)
scaler_pca_kernel(const float* __restrict__ feat, int n_cells, const int32_t* __restrict__ n_dev,
                  int F, int C, const double* __restrict__ center, const double* __restrict__ scale,
                  int center_is_f32, const double* __restrict__ comp_t,
                  const double* __restrict__ offset, int f32_flow, double* __restrict__ z_out) {
    extern __shared__ __align__(16) unsigned char pca_smem[];
    double (*ws)[PCT] = reinterpret_cast<double (*)[PCT]>(pca_smem);
    double (*xs)[XPAD] = reinterpret_cast<double (*)[XPAD]>(pca_smem + sizeof(double) * PFT * PCT);
    const int n = dev_count(n_cells, n_dev);
    const int cell0 = blockIdx.x * PB;
    if (cell0 >= n) return;
    const int c_tile0 = blockIdx.y * PCT;
    const int tid = threadIdx.x, cx = tid & 31, cy = tid >> 5;
    double acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

    for (int f0 = 0; f0 < F; f0 += PFT) {
        __syncthreads();
        for (int idx = tid; idx < PB * PFT; idx += PT) {
            const int i = idx / PFT, ff = idx - i * PFT;
            const int f = f0 + ff, cell = cell0 + i;
            float v = 0.f;
            if (cell < n && f < F) {
                v = __ldg(feat + (size_t)cell * F + f);
                if (center) {
                    if (center_is_f32) v = __fsub_rn(v, (float)center[f]);
                    else v = (float)__dsub_rn((double)v, center[f]);
                }
                if (scale) v = (float)__ddiv_rn((double)v, scale[f]);
            }
            xs[ff][i] = (double)v;
        }
        for (int idx = tid; idx < PFT * PCT; idx += PT) {
            const int ff = idx / PCT, c = idx - ff * PCT;
            const int f = f0 + ff, cc = c_tile0 + c;
            ws[ff][c] = (f < F && cc < C) ? __ldg(comp_t + (size_t)f * C + cc) : 0.0;
        }
        __syncthreads();
#pragma unroll 4
        for (int ff = 0; ff < PFT; ++ff) {
            double w[4], x[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = ws[ff][cx + 32 * j];
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                const double2 t = *reinterpret_cast<const double2*>(&xs[ff][cy * 8 + i]);
                x[i] = t.x; x[i + 1] = t.y;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(x[i], w[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = c_tile0 + cx + 32 * j;
        if (c >= C) continue;
        const double off = offset[c];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int cell = cell0 + cy * 8 + i;
            if (cell < n) {
                double z;
                if (f32_flow) z = (double)__fsub_rn((float)acc[i][j], (float)off);
                else z = __dsub_rn(acc[i][j], off);
                z_out[(size_t)cell * C + c] = z;
            }
        }
    }
}

// ---- fp64 tensor-core form of the projection (mma.sync m8n8k4 f64), software pipelined ----
// Measured on this B200 (profiles/fp64_peak_test.cu): DFMA 36 TFLOP/s, DMMA 37 TFLOP/s; the
// register-tiled kernel above reaches ~9 because every 32-feature stage first waits for its
// global loads.  Here the components come through cp.async into a second buffer and the raw
// features of the next stage are prefetched into registers while the current stage computes.
// Block = 16 warps = 64 cells x 104 components (13 n8 tiles; blockIdx.y tiles wider PCAs);
// warp = 16 cells x 3-4 n8 tiles (4 cell groups x 4 component groups).  A call scores ~15k cells =
// 236 blocks on 148 SMs, so the warps have to come from inside the block: with 4 warps per SM the
// DMMA pipe was 21 % active (ncu), latency bound.  Shared-memory pitches put the 64-bit
// fragment loads of a half-warp on 16 distinct bank pairs (x: 36 doubles per cell row, w: 108
// per feature row).  comp_pad is the [F rounded to 32][C rounded to 104] zero-padded copy.
constexpr int DT = 512;          // threads
constexpr int DM = 64;           // cells per block
constexpr int DK = 32;           // features per stage
constexpr int DN = 104;          // components per block tile
constexpr int DXP = DK + 4;      // x pitch (doubles)
constexpr int DWP = DN + 4;      // w pitch (doubles)
constexpr int DXR = DM / (DT / 32);   // cell rows per warp in the x tile

__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(DT, 2)
scaler_pca_dmma_kernel(const float* __restrict__ feat, int n_cells, const int32_t* __restrict__ n_dev,
                       int F, int C, int CP, const double* __restrict__ center, const double* __restrict__ scale,
                       int center_is_f32, const double* __restrict__ comp_pad,
                       const double* __restrict__ offset, int f32_flow, double* __restrict__ z_out) {
    extern __shared__ __align__(16) unsigned char pca_smem[];
    double* xs = reinterpret_cast<double*>(pca_smem);                 // [2][DM][DXP]
    double* ws = xs + 2 * DM * DXP;                                   // [2][DK][DWP]
    const int n = dev_count(n_cells, n_dev);
    const int cell0 = blockIdx.x * DM;
    if (cell0 >= n) return;
    const int c_tile0 = blockIdx.y * DN;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gid = lane >> 2, tig = lane & 3;
    const int n_stages = (F + DK - 1) / DK;
    // component group of this warp: n8 tiles [nt0, nt0 + ntc) = {0-3, 4-6, 7-9, 10-12}
    const int mg = warp & 3, ng = warp >> 2;
    const int nt0 = ng == 0 ? 0 : 1 + 3 * ng, ntc = ng == 0 ? 4 : 3;
    static_assert(DN / 8 == 13 && DT == 512, "warp tiling below assumes 13 n8 tiles over 4 component groups");
    double acc[2][4][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

    auto issue_w = [&](int stage, int buf) {          // DK x DN doubles as 16-byte chunks
        const double* src = comp_pad + (size_t)stage * DK * CP + c_tile0;
        double* dst = ws + buf * DK * DWP;
        for (int idx = tid; idx < DK * (DN / 2); idx += DT) {
            const int ff = idx / (DN / 2), c2 = idx - ff * (DN / 2);
            const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + ff * DWP + 2 * c2);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + (size_t)ff * CP + 2 * c2) : "memory");
        }
    };
    float xr[DXR];
    auto load_x = [&](int stage) {                    // raw features: one cell row per warp iteration
        const int f = stage * DK + lane;
#pragma unroll
        for (int j = 0; j < DXR; ++j) {
            const int cell = cell0 + warp + j * (DT / 32);
            xr[j] = (cell < n && f < F) ? __ldg(feat + (size_t)cell * F + f) : 0.f;
        }
    };
    auto store_x = [&](int stage, int buf) {          // scaler applied on the way (sklearn's float32 flow)
        const int f = stage * DK + lane;
        double cen = 0.0, sc = 1.0;
        if (f < F) {
            if (center) cen = center[f];
            if (scale) sc = scale[f];
        }
        double* dst = xs + buf * DM * DXP;
#pragma unroll
        for (int j = 0; j < DXR; ++j) {
            float v = xr[j];
            if (f < F) {
                if (center) {
                    if (center_is_f32) v = __fsub_rn(v, (float)cen);
                    else v = (float)__dsub_rn((double)v, cen);
                }
                if (scale) v = (float)__ddiv_rn((double)v, sc);
            }
            dst[(warp + j * (DT / 32)) * DXP + lane] = (double)v;
        }
    };

    issue_w(0, 0);
    load_x(0);
    store_x(0, 0);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    for (int st = 0; st < n_stages; ++st) {
        const int buf = st & 1;
        const bool more = st + 1 < n_stages;
        if (more) {
            issue_w(st + 1, buf ^ 1);
            load_x(st + 1);
        }
        const double* xb = xs + buf * DM * DXP;
        const double* wb = ws + buf * DK * DWP;
#pragma unroll
        for (int k4 = 0; k4 < DK / 4; ++k4) {
            const double a0 = xb[(mg * 16 + gid) * DXP + k4 * 4 + tig];
            const double a1 = xb[(mg * 16 + 8 + gid) * DXP + k4 * 4 + tig];
            const double* wrow = wb + (k4 * 4 + tig) * DWP + nt0 * 8 + gid;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                if (nt < ntc) {
                    const double b = wrow[nt * 8];
                    dmma884(acc[0][nt], a0, b);
                    dmma884(acc[1][nt], a1, b);
                }
            }
        }
        if (more) {
            store_x(st + 1, buf ^ 1);
            asm volatile("cp.async.wait_all;" ::: "memory");
        }
        __syncthreads();
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        const int cell = cell0 + mg * 16 + mt * 8 + gid;
        if (cell >= n) continue;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int c = c_tile0 + (nt0 + nt) * 8 + 2 * tig + j;
                if (nt >= ntc || c >= C) continue;
                const double off = offset[c];
                double z;
                if (f32_flow) z = (double)__fsub_rn((float)acc[mt][nt][j], (float)off);
                else z = __dsub_rn(acc[mt][nt][j], off);
                z_out[(size_t)cell * C + c] = z;
            }
    }
}

constexpr int SCELLS = 8;   // cells per block in the SVM kernel

// dec = sum_i coef_i * exp(-gamma * ||z - sv_i||^2) - rho  (fp64, direct difference).
__global__ void __launch_bounds__(PT)
svm_rbf_kernel(const double* __restrict__ z, int n_cells, const int32_t* __restrict__ n_dev, int D,
               const double* __restrict__ sv_t, const double* __restrict__ coef, int n_sv,
               int n_sv_pad, double gamma, double rho, double* __restrict__ dec,
               int8_t* __restrict__ pred) {
    extern __shared__ double zs[];  // [SCELLS][D]
    __shared__ double red[SCELLS][PT / 32];
    const int n = dev_count(n_cells, n_dev);
    const int cell0 = blockIdx.x * SCELLS;
    if (cell0 >= n) return;
    const int nc = min(SCELLS, n - cell0);
    for (int i = threadIdx.x; i < SCELLS * D; i += PT) {
        const int k = i / D;
        zs[i] = k < nc ? z[(size_t)cell0 * D + i] : 0.0;
    }
    __syncthreads();
    double part[SCELLS];
#pragma unroll
    for (int k = 0; k < SCELLS; ++k) part[k] = 0.0;
    for (int i = threadIdx.x; i < n_sv; i += PT) {
        double d2[SCELLS];
#pragma unroll
        for (int k = 0; k < SCELLS; ++k) d2[k] = 0.0;
        for (int d = 0; d < D; ++d) {
            const double s = __ldg(sv_t + (size_t)d * n_sv_pad + i);
#pragma unroll
            for (int k = 0; k < SCELLS; ++k) {
                const double df = zs[k * D + d] - s;
                d2[k] = fma(df, df, d2[k]);
            }
        }
        const double a = coef[i];
#pragma unroll
        for (int k = 0; k < SCELLS; ++k) part[k] = fma(a, exp(-gamma * d2[k]), part[k]);
    }
#pragma unroll
    for (int k = 0; k < SCELLS; ++k) {
        const double v = warp_sum(part[k]);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < nc) {
        double s = 0.0;
        for (int w = 0; w < PT / 32; ++w) s += red[threadIdx.x][w];
        s -= rho;
        dec[cell0 + threadIdx.x] = s;
        pred[cell0 + threadIdx.x] = s > 0.0 ? 1 : -1;     // svm.cpp:2841
    }
}

// K6: acc[strain] += {1, cons anomalous, mod anomalous, mse, mse^2, mae, mae^2, 0}
__global__ void __launch_bounds__(256)
strain_accumulate_kernel(const cia_cell* __restrict__ cells, int n_cells,
                         const int32_t* __restrict__ n_dev, const float* __restrict__ mse,
                         const float* __restrict__ mae, const int8_t* __restrict__ pc,
                         const int8_t* __restrict__ pm, const int32_t* __restrict__ field_strain,
                         double* __restrict__ acc, int n_strains) {
    const int n = dev_count(n_cells, n_dev);
    const int lane = threadIdx.x & 31;
    for (int base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < n; base += gridDim.x * blockDim.x) {
        const int i = base + lane;
        const bool valid = i < n;
        int strain = -1;
        double v[7] = {0, 0, 0, 0, 0, 0, 0};
        if (valid) {
            strain = field_strain ? field_strain[cells[i].field] : 0;
            if (strain < 0 || strain >= n_strains) strain = -1;
            const double a = (double)mse[i], b = (double)mae[i];
            v[0] = 1.0; v[1] = pc[i] == -1 ? 1.0 : 0.0; v[2] = pm[i] == -1 ? 1.0 : 0.0;
            v[3] = a; v[4] = a * a; v[5] = b; v[6] = b * b;
        }
        const unsigned act = __ballot_sync(0xffffffffu, strain >= 0);
        if (strain >= 0) {
            const unsigned m = __match_any_sync(act, strain);
            const int leader = __ffs(m) - 1;
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                // fixed-order reduction over the group's lanes
                double s = 0.0;
                for (unsigned mm = m; mm; mm &= mm - 1) {
                    const int src = __ffs(mm) - 1;
                    s += __shfl_sync(m, v[k], src);
                }
                if (lane == leader) atomicAdd(&acc[(size_t)strain * 8 + k], s);
            }
        }
    }
}

}  // namespace

int k_svm_decision(cia_ctx* h, const float* features, int n, const int32_t* n_dev,
                   double* dec_cons, double* dec_mod, int8_t* pred_cons, int8_t* pred_mod,
                   double* pca_out, cudaStream_t s) {
    if (n <= 0) return CIA_OK;
    if (!h->sp.loaded || !h->svm[0].loaded || !h->svm[1].loaded) {
        h->err = "cia_svm_decision: scaler/pca/svm artifacts not loaded";
        return CIA_E_STATE;
    }
    const ScalerPca& sp = h->sp;
    double* z = pca_out;
    if (!z) {
        int rc = ws_reserve(h, h->ws_feat, (size_t)n * sp.C * sizeof(double));
        if (rc) return rc;
        z = (double*)h->ws_feat.p;
    }
    if (first_use(h, (const void*)svm_rbf_kernel)) {
        CIA_CUDA(cudaFuncSetAttribute(svm_rbf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CIA_CUDA(cudaFuncSetAttribute(scaler_pca_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        CIA_CUDA(cudaFuncSetAttribute(scaler_pca_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    }
    static const bool vector_pca = getenv("CIA_PCA_VECTOR") != nullptr;   // A/B switch: CUDA-core fp64 tile
    if (vector_pca) {
        const size_t sm1 = sizeof(double) * PFT * (PCT + XPAD);
        scaler_pca_kernel<<<dim3((n + PB - 1) / PB, (sp.C + PCT - 1) / PCT), PT, sm1, s>>>(
            features, n, n_dev, sp.F, sp.C, sp.has_center ? sp.center : nullptr,
            sp.has_scale ? sp.scale : nullptr, sp.center_is_f32, sp.comp_t, sp.offset, sp.f32_flow, z);
    } else {
        const size_t sm1 = sizeof(double) * 2 * (DM * DXP + DK * DWP);
        scaler_pca_dmma_kernel<<<dim3((n + DM - 1) / DM, sp.CP / DN), DT, sm1, s>>>(
            features, n, n_dev, sp.F, sp.C, sp.CP, sp.has_center ? sp.center : nullptr,
            sp.has_scale ? sp.scale : nullptr, sp.center_is_f32, sp.comp_pad, sp.offset, sp.f32_flow, z);
    }
    CIA_LAUNCH_CHECK();
    for (int which = 0; which < 2; ++which) {
        const SvmModel& m = h->svm[which];
        if (m.dim != sp.C) { h->err = "svm dimension != pca components"; return CIA_E_STATE; }
        const size_t sm2 = (size_t)SCELLS * m.dim * sizeof(double);
        svm_rbf_kernel<<<(n + SCELLS - 1) / SCELLS, PT, sm2, s>>>(
            z, n, n_dev, m.dim, m.sv_t, m.coef, m.n_sv, m.n_sv_pad, m.gamma, m.rho,
            which == 0 ? dec_cons : dec_mod, which == 0 ? pred_cons : pred_mod);
        CIA_LAUNCH_CHECK();
    }
    return CIA_OK;
}

int k_strain_accumulate(cia_ctx* h, const cia_cell* cells, int n, const int32_t* n_dev,
                        const cia_scores* sc, const int32_t* field_strain, double* acc,
                        int n_strains, cudaStream_t s) {
    if (n <= 0 || !acc) return CIA_OK;
    int blocks = (n + 255) / 256;
    if (blocks > h->num_sms * 8) blocks = h->num_sms * 8;
    strain_accumulate_kernel<<<blocks, 256, 0, s>>>(cells, n, n_dev, sc->mse, sc->mae,
                                                    sc->pred_conservative, sc->pred_moderate,
                                                    field_strain, acc, n_strains);
    CIA_LAUNCH_CHECK();
    return CIA_OK;
}
