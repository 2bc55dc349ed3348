// cae_fp32.cu -- K3, fp32 CUDA-core path of the convolutional autoencoder forward
// pass with fused ReLU + BatchNorm + max-pool epilogues, nearest up-sampling folded
// into the next layer's gather, and sigmoid + squared / absolute error reduction fused
// into the last layer.
//
// Replaces autoencoder.predict + the MSE / MAE of improved_detection.py:125-127 and
// encoder.predict + flatten of :130-131; architecture CAE_improved_modeltrain.py:188-216.
// This is the exact-fp32 path (precision = 0): the parity anchor for the tensor-core
// path in cae_tc.cu and the producer of fp32-faithful encoder features.
#include "common.cuh"

namespace {

constexpr int CT = 256;   // threads per block

// One block: an 8 x TSX tile of conv pixels (TSX = min(16, R)) x all COUT channels of one
// cell.  thread -> (channel group cg = warp id, 2x2 pixel unit = lane).  Products are
// accumulated in fp32 over FLUSH input channels (9*FLUSH FMAs), then flushed into an
// fp64 accumulator, so a layer output is within ~2e-7 relative of the correctly rounded
// fp32 sum regardless of K -- the one-class SVM downstream amplifies feature noise
// (1e-6 relative -> 7e-5 in the decision value), see DESIGN.md "precision".
template <int CIN, int COUT, int R, bool POOL, bool UPS>
__global__ void __launch_bounds__(CT)
conv3x3_kernel(const float* __restrict__ in, float* __restrict__ out,
               const float* __restrict__ wgt,    // [9][CIN][COUT]
               const float* __restrict__ bias, const float* __restrict__ bn_s,
               const float* __restrict__ bn_t, int n_cells, const int32_t* __restrict__ n_dev,
               int cell0) {
    constexpr int TSX = R < 16 ? R : 16;
    constexpr int TSY = 8;
    constexpr int TX = R / TSX, TY = R / TSY;
    constexpr int TILES = TX * TY;
    constexpr int CK = CIN < 8 ? CIN : 8;
    constexpr int FLUSH = CK < 2 ? CK : 2;
    constexpr int CPT = COUT / 8;
    constexpr int PPX = TSX + 2, PPY = TSY + 2;
    constexpr int RIN = UPS ? R / 2 : R;
    constexpr int UX = TSX / 2;
    constexpr int UNITS = UX * (TSY / 2);

    __shared__ __align__(16) float in_s[CK][PPY][PPX + 1];
    __shared__ __align__(16) float w_s[9][CK][COUT];

    const int n = dev_count(n_cells, n_dev);
    const int cell = cell0 + blockIdx.x / TILES;
    if (cell >= n) return;
    const int tile = blockIdx.x % TILES;
    const int ty0 = (tile / TX) * TSY, tx0 = (tile % TX) * TSX;
    const int tid = threadIdx.x;
    const int cg = tid >> 5, unit = tid & 31;
    const bool active = unit < UNITS;
    const int uy = unit / UX, ux = unit % UX;

    float acc[4][CPT];
    double dacc[4][CPT];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int k = 0; k < CPT; ++k) { acc[p][k] = 0.f; dacc[p][k] = 0.0; }

    const float* src = in + (size_t)cell * RIN * RIN * CIN;

    for (int c0 = 0; c0 < CIN; c0 += CK) {
        __syncthreads();
        // stage the input patch (zero padded; nearest up-sampling folded in)
        for (int i = tid; i < PPY * PPX * CK; i += CT) {
            const int c = i % CK, pix = i / CK;
            const int py = pix / PPX, px = pix % PPX;
            const int y = ty0 + py - 1, x = tx0 + px - 1;
            float v = 0.f;
            if (y >= 0 && y < R && x >= 0 && x < R) {
                const int ys = UPS ? (y >> 1) : y, xs = UPS ? (x >> 1) : x;
                v = __ldg(src + ((size_t)ys * RIN + xs) * CIN + c0 + c);
            }
            in_s[c][py][px] = v;
        }
        for (int i = tid; i < 9 * CK * COUT; i += CT) {
            const int co = i % COUT, r = i / COUT;
            const int c = r % CK, tap = r / CK;
            w_s[tap][c][co] = __ldg(wgt + ((size_t)tap * CIN + c0 + c) * COUT + co);
        }
        __syncthreads();
        if (active) {
#pragma unroll 1
            for (int cc = 0; cc < CK; cc += FLUSH) {
#pragma unroll
                for (int c = cc; c < cc + FLUSH; ++c) {
                    float win[4][4];
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) win[a][b] = in_s[c][2 * uy + a][2 * ux + b];
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const float* wp = &w_s[dy * 3 + dx][c][cg * CPT];
#pragma unroll
                            for (int k = 0; k < CPT; ++k) {
                                const float w = wp[k];
                                acc[0][k] = fmaf(win[dy][dx], w, acc[0][k]);
                                acc[1][k] = fmaf(win[dy][dx + 1], w, acc[1][k]);
                                acc[2][k] = fmaf(win[dy + 1][dx], w, acc[2][k]);
                                acc[3][k] = fmaf(win[dy + 1][dx + 1], w, acc[3][k]);
                            }
                        }
                }
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int k = 0; k < CPT; ++k) { dacc[p][k] += (double)acc[p][k]; acc[p][k] = 0.f; }
            }
        }
    }
    if (!active) return;

    // epilogue: round to fp32 -> bias -> ReLU -> BN affine -> (2x2 max | store)
    const int cbase = cg * CPT;
    if (POOL) {
        constexpr int RO = R / 2;
        float* dst = out + (((size_t)cell * RO + (ty0 / 2 + uy)) * RO + (tx0 / 2 + ux)) * COUT + cbase;
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            const float b = bias[cbase + k], s = bn_s[cbase + k], t = bn_t[cbase + k];
            float m = -INFINITY;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                float v = __fadd_rn((float)dacc[p][k], b);
                v = v > 0.f ? v : 0.f;
                v = __fadd_rn(__fmul_rn(v, s), t);
                m = fmaxf(m, v);
            }
            dst[k] = m;
        }
    } else {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int y = ty0 + 2 * uy + (p >> 1), x = tx0 + 2 * ux + (p & 1);
            float* dst = out + (((size_t)cell * R + y) * R + x) * COUT + cbase;
#pragma unroll
            for (int k = 0; k < CPT; ++k) {
                float v = __fadd_rn((float)dacc[p][k], bias[cbase + k]);
                v = v > 0.f ? v : 0.f;
                dst[k] = __fadd_rn(__fmul_rn(v, bn_s[cbase + k]), bn_t[cbase + k]);
            }
        }
    }
}

// Last layer: conv3x3 32 -> 1 on the up-sampled 32x32x32 activation, sigmoid, and the
// per-cell mean squared / absolute error against the input crop.  One block per cell.
__global__ void __launch_bounds__(CT)
final_layer_kernel(const float* __restrict__ a6, const float* __restrict__ crops,
                   const float* __restrict__ wgt /* [9][32] */, const float* __restrict__ bias,
                   float* __restrict__ mse, float* __restrict__ mae, float* __restrict__ recon,
                   int n_cells, const int32_t* __restrict__ n_dev, int cell0) {
    constexpr int CIN = 32, RL = 32;
    __shared__ float in_s[CIN][10][11];
    __shared__ float w_s[9][CIN];
    __shared__ float red_s[2][CT / 32];
    const int n = dev_count(n_cells, n_dev);
    const int cell = cell0 + blockIdx.x;
    if (cell >= n) return;
    const int tid = threadIdx.x;
    for (int i = tid; i < 9 * CIN; i += CT) w_s[i / CIN][i % CIN] = wgt[i];
    const float b = bias[0];
    const float* src = a6 + (size_t)cell * RL * RL * CIN;
    const float* x = crops + (size_t)cell * 4096;
    float se = 0.f, ae = 0.f;
    for (int tile = 0; tile < 16; ++tile) {
        const int ty0 = (tile >> 2) * 16, tx0 = (tile & 3) * 16;
        const int ly0 = ty0 / 2 - 1, lx0 = tx0 / 2 - 1;
        __syncthreads();
        for (int i = tid; i < 100 * CIN; i += CT) {
            const int c = i % CIN, pix = i / CIN;
            const int py = pix / 10, px = pix % 10;
            const int ys = ly0 + py, xs = lx0 + px;
            float v = 0.f;
            if (ys >= 0 && ys < RL && xs >= 0 && xs < RL) v = __ldg(src + ((size_t)ys * RL + xs) * CIN + c);
            in_s[c][py][px] = v;
        }
        __syncthreads();
        const int oy = ty0 + (tid >> 4), ox = tx0 + (tid & 15);
        float acc = 0.f;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int y = oy + dy - 1, xx = ox + dx - 1;
                if (y < 0 || y >= 64 || xx < 0 || xx >= 64) continue;
                const int py = (y >> 1) - ly0, px = (xx >> 1) - lx0;
#pragma unroll 8
                for (int c = 0; c < CIN; ++c) acc = fmaf(in_s[c][py][px], w_s[dy * 3 + dx][c], acc);
            }
        const float v = acc + b;
        const float y = 1.f / (1.f + expf(-v));
        const float xv = x[oy * 64 + ox];
        const float d = xv - y;
        se += d * d;
        ae += fabsf(d);
        if (recon) recon[(size_t)cell * 4096 + oy * 64 + ox] = y;
    }
    se = warp_sum(se);
    ae = warp_sum(ae);
    if ((tid & 31) == 0) { red_s[0][tid >> 5] = se; red_s[1][tid >> 5] = ae; }
    __syncthreads();
    if (tid == 0) {
        float s = 0.f, a = 0.f;
        for (int i = 0; i < CT / 32; ++i) { s += red_s[0][i]; a += red_s[1][i]; }
        mse[cell] = s / 4096.f;
        mae[cell] = a / 4096.f;
    }
}

template <int CIN, int COUT, int R, bool POOL, bool UPS>
int launch_conv(cia_ctx* h, const CaeWeights& w, int layer, const float* in, float* out, int n,
                const int32_t* n_dev, int cell0, int chunk, cudaStream_t s) {
    constexpr int TSX = R < 16 ? R : 16;
    constexpr int TILES = (R / TSX) * (R / 8);
    conv3x3_kernel<CIN, COUT, R, POOL, UPS><<<chunk * TILES, CT, 0, s>>>(
        in, out, w.kernel[layer], w.bias[layer], w.bn_scale[layer], w.bn_shift[layer], n, n_dev, cell0);
    CIA_LAUNCH_CHECK();
    return CIA_OK;
}

}  // namespace

// Third encoder layer alone in exact fp32 (input: fp32 NHWC 16x16x64, output: features).
int k_conv3_fp32(cia_ctx* h, const CaeWeights& w, const float* a2, int n, const int32_t* n_dev,
                 float* features, int cell0, int chunk, cudaStream_t s) {
    return launch_conv<64, 32, 16, true, false>(h, w, 2, a2, features, n, n_dev, cell0, chunk, s);
}

// encoder.predict + flatten (det:130-131) alone: the three encoder layers in exact fp32.
int k_encoder_fp32(cia_ctx* h, const CaeWeights& w, const float* crops, int n, const int32_t* n_dev,
                   float* features, cudaStream_t s) {
    if (n <= 0) return CIA_OK;
    if (!w.loaded || w.n_conv < 3) { h->err = "encoder weights not loaded"; return CIA_E_STATE; }
    const int CH = 2048;
    const size_t a1 = 32 * 32 * 32, a2 = 16 * 16 * 64;
    int rc = ws_reserve(h, h->ws_act, (size_t)CH * (a1 + a2) * sizeof(float));
    if (rc) return rc;
    float* A1 = (float*)h->ws_act.p;
    float* A2 = A1 + CH * a1;
    for (int c0 = 0; c0 < n; c0 += CH) {
        const int chunk = (n - c0) < CH ? (n - c0) : CH;
        float* a1p = A1 - (size_t)c0 * a1; float* a2p = A2 - (size_t)c0 * a2;
        if ((rc = launch_conv<1, 32, 64, true, false>(h, w, 0, crops, a1p, n, n_dev, c0, chunk, s))) return rc;
        if ((rc = launch_conv<32, 64, 32, true, false>(h, w, 1, a1p, a2p, n, n_dev, c0, chunk, s))) return rc;
        if ((rc = launch_conv<64, 32, 16, true, false>(h, w, 2, a2p, features, n, n_dev, c0, chunk, s))) return rc;
    }
    return CIA_OK;
}

// Activations per cell (fp32): A1 32x32x32, A2 16x16x64, A3 8x8x32, A4 8x8x32, A5 16x16x64,
// A6 32x32x32.  Buffers are indexed by absolute cell so chunks never alias.
int k_cae_forward_fp32(cia_ctx* h, const float* crops, int n, const int32_t* n_dev, float* mse,
                       float* mae, float* features, cudaStream_t s) {
    if (n <= 0) return CIA_OK;
    const CaeWeights& ae = h->cae[0];
    if (!ae.loaded || ae.n_conv != 7) { h->err = "cia_cae_forward: autoencoder not loaded"; return CIA_E_STATE; }
    const int CH = 2048;   // cells per chunk
    const size_t a1 = 32 * 32 * 32, a2 = 16 * 16 * 64, a3 = 8 * 8 * 32;
    const size_t per_cell = a1 + a2 + a3 + a3 + a2 + a1;
    int rc = ws_reserve(h, h->ws_act, (size_t)CH * per_cell * sizeof(float));
    if (rc) return rc;
    float* A1 = (float*)h->ws_act.p;
    float* A2 = A1 + CH * a1;
    float* A3 = A2 + CH * a2;
    float* A4 = A3 + CH * a3;
    float* A5 = A4 + CH * a3;
    float* A6 = A5 + CH * a2;
    const bool sep = h->cae[1].loaded;
    for (int c0 = 0; c0 < n; c0 += CH) {
        const int chunk = (n - c0) < CH ? (n - c0) : CH;
        // workspace buffers are chunk-relative: shift pointers so absolute cell indexing works
        float* a1p = A1 - (size_t)c0 * a1; float* a2p = A2 - (size_t)c0 * a2;
        float* a3p = A3 - (size_t)c0 * a3; float* a4p = A4 - (size_t)c0 * a3;
        float* a5p = A5 - (size_t)c0 * a2; float* a6p = A6 - (size_t)c0 * a1;
        float* feat_ae = (features && !sep) ? features : a3p;
        if ((rc = launch_conv<1, 32, 64, true, false>(h, ae, 0, crops, a1p, n, n_dev, c0, chunk, s))) return rc;
        if ((rc = launch_conv<32, 64, 32, true, false>(h, ae, 1, a1p, a2p, n, n_dev, c0, chunk, s))) return rc;
        if ((rc = launch_conv<64, 32, 16, true, false>(h, ae, 2, a2p, feat_ae, n, n_dev, c0, chunk, s))) return rc;
        if ((rc = launch_conv<32, 32, 8, false, false>(h, ae, 3, feat_ae, a4p, n, n_dev, c0, chunk, s))) return rc;
        if ((rc = launch_conv<32, 64, 16, false, true>(h, ae, 4, a4p, a5p, n, n_dev, c0, chunk, s))) return rc;
        if ((rc = launch_conv<64, 32, 32, false, true>(h, ae, 5, a5p, a6p, n, n_dev, c0, chunk, s))) return rc;
        final_layer_kernel<<<chunk, CT, 0, s>>>(a6p, crops, ae.kernel[6], ae.bias[6], mse, mae, nullptr,
                                                n, n_dev, c0);
        CIA_LAUNCH_CHECK();
        if (features && sep) {
            // encoder.keras holds different weights (det:29, 130): second encoder pass
            const CaeWeights& en = h->cae[1];
            if ((rc = launch_conv<1, 32, 64, true, false>(h, en, 0, crops, a1p, n, n_dev, c0, chunk, s))) return rc;
            if ((rc = launch_conv<32, 64, 32, true, false>(h, en, 1, a1p, a2p, n, n_dev, c0, chunk, s))) return rc;
            if ((rc = launch_conv<64, 32, 16, true, false>(h, en, 2, a2p, features, n, n_dev, c0, chunk, s))) return rc;
        }
    }
    return CIA_OK;
}
