// crop.cu -- K2: bbox crop + CLAHE + anti-aliased resize to 64x64 + float32 cast,
// one CTA per cell, everything between the uint16 bbox read and the 64x64 write
// staying in shared memory.
//
// Replaces improved_detection.py:88 (crop), :98 exposure.equalize_adapthist(cell,
// clip_limit=0.02), :99 resize(..., (64, 64), anti_aliasing=True) and the float32
// cast of :122 (training twin CAE_improved_modeltrain.py:80, 92-93, 332).
// The operation order of scikit-image's CLAHE is replicated step for step
// (SURVEY.md A.2): the integer core (14-bit quantise, 256-bin tile histograms,
// clip + serial redistribution, cumulative mapping) is exact, and the float steps use
// the same precision and rounding (fp64 products rounded to fp32 and accumulated in
// fp32, truncation to uint16), so the uint16 levels are bit-identical.  The resize
// (A.3) runs in fp64 and is rounded to fp32 once, like det:122.
//
// Algorithmic HBM traffic per cell: 2*h*w bytes read + 16 KiB written.
#include "common.cuh"

namespace {

constexpr int K2_THREADS = 512;
constexpr int K2_WARPS = K2_THREADS / 32;
constexpr int NBINS = 256;
constexpr int BIN_SIZE = 65;           // 1 + 16384 / 256
constexpr int MAX_RADIUS = 31;         // gaussian radius for bbox sides <= 1024
constexpr int MAX_SIDE = 1024;

struct Geom {
    int h, w, kh, kw, ntr, ntc, npix, clim;
};

__device__ __forceinline__ int reflect_idx(int i, int n) {   // numpy 'reflect'
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}
__device__ __forceinline__ int mirror_idx(int i, int n) {    // ndimage 'mirror'
    if (n == 1) return 0;
    const int p = 2 * n - 2;
    i = i < 0 ? -i : i;
    i %= p;
    return i >= n ? p - i : i;
}

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }
__host__ __device__ inline size_t cell_bytes(int h, int w, int ntr, int ntc) {
    const size_t hw = (size_t)h * w;
    size_t maps = (size_t)ntr * ntc * NBINS * sizeof(uint16_t);
    size_t t = (size_t)CIA_CROP * w * sizeof(double);
    return align16(hw) + align16(2 * hw) + align16(maps > t ? maps : t);
}

// One warp: clip + redistribute a 256-bin histogram held 8 bins per lane
// (bin = lane + 32*m), then cumulative mapping.  skimage clip_histogram / map_histogram.
__device__ __forceinline__ void clip_and_map(int (&hv)[8], int clim, int npix, int lane,
                                             uint16_t* __restrict__ map_out) {
    int ex = 0;
#pragma unroll
    for (int m = 0; m < 8; ++m)
        if (hv[m] > clim) { ex += hv[m] - clim; hv[m] = clim; }
    ex = warp_sum(ex);
    const int incr = ex / NBINS;
    const int upper = clim - incr;
    int nlow = 0;
#pragma unroll
    for (int m = 0; m < 8; ++m)
        if (hv[m] < upper) { hv[m] += incr; ++nlow; }
    ex -= warp_sum(nlow) * incr;
    int midsum = 0, nmid = 0;
#pragma unroll
    for (int m = 0; m < 8; ++m)
        if (hv[m] >= upper && hv[m] < clim) { midsum += hv[m]; ++nmid; hv[m] = clim; }
    ex += warp_sum(midsum) - warp_sum(nmid) * clim;

    while (ex > 0) {
        const int prev = ex;
        bool stuck = false;
        for (int index = 0; index < NBINS; ++index) {
            int cnt = 0;
#pragma unroll
            for (int m = 0; m < 8; ++m) cnt += __popc(__ballot_sync(0xffffffffu, hv[m] < clim));
            if (cnt == 0) { stuck = true; break; }   // nothing can move any more
            int step = cnt / ex;
            if (step < 1) step = 1;
            int moved = 0;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int b = lane + 32 * m;
                const bool sel = b >= index && hv[m] < clim && ((b - index) % step) == 0;
                if (sel) ++hv[m];
                moved += __popc(__ballot_sync(0xffffffffu, sel));
            }
            ex -= moved;
            if (ex <= 0) break;
        }
        if (stuck || prev == ex) break;
    }

    // cumulative sum in bin order, scale, clip, truncate
    const double scale = __ddiv_rn(16383.0, (double)npix);
    int carry = 0;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        int v = hv[m];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        const int incl = v + carry;
        carry += __shfl_sync(0xffffffffu, v, 31);
        double mv = __dmul_rn((double)incl, scale);
        if (mv > 16383.0) mv = 16383.0;
        map_out[lane + 32 * m] = (uint16_t)(int)mv;
    }
}

__device__ __forceinline__ void block_minmax(int& mn, int& mx, int* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) { sh[wid] = mn; sh[K2_WARPS + wid] = mx; }
    __syncthreads();
    mn = sh[0]; mx = sh[K2_WARPS];
#pragma unroll
    for (int i = 1; i < K2_WARPS; ++i) { mn = min(mn, sh[i]); mx = max(mx, sh[K2_WARPS + i]); }
}

// cls: 0/1 = shared-memory classes (dynamic smem = smem_bytes), 2 = global-scratch class.
__global__ void __launch_bounds__(K2_THREADS)
crop_clahe_resize_kernel(const uint16_t* __restrict__ images, int H, int W,
                         const cia_cell* __restrict__ cells, int n_cells,
                         const int32_t* __restrict__ n_cells_dev, double clip_limit,
                         float* __restrict__ crops32, double* __restrict__ crops64,
                         int cls, size_t lo_bytes, size_t hi_bytes,
                         unsigned char* __restrict__ gscratch, size_t gscratch_per_cta,
                         int32_t* status, uint16_t* __restrict__ levels_out,
                         const int64_t* __restrict__ level_offsets) {
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ uint32_t hist_s[K2_WARPS][NBINS];
    __shared__ int red_s[2 * K2_WARPS];
    __shared__ double gw_s[2][2 * MAX_RADIUS + 2];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = dev_count(n_cells, n_cells_dev);

    for (int cell = blockIdx.x; cell < n; cell += gridDim.x) {
        const cia_cell C = cells[cell];
        Geom g;
        g.h = C.maxr - C.minr; g.w = C.maxc - C.minc;
        g.kh = max(g.h / 8, 1); g.kw = max(g.w / 8, 1);
        g.ntr = (g.h + g.kh - 1) / g.kh; g.ntc = (g.w + g.kw - 1) / g.kw;
        g.npix = g.kh * g.kw;
        {
            double cl = __dmul_rn(clip_limit, (double)g.npix);
            if (!(cl >= 1.0)) cl = 1.0;
            g.clim = clip_limit > 0.0 ? (int)cl : 16384;
        }
        const size_t need = cell_bytes(g.h, g.w, g.ntr, g.ntc);
        int my_cls;
        if (g.h > MAX_SIDE || g.w > MAX_SIDE) {
            if (cls == 2 && tid == 0) raise_status(status, CIA_E_UNSUPPORTED);
            if (cls == 2) {
                // defined output for the unsupported cell: zeros
                for (int i = tid; i < CIA_CROP * CIA_CROP; i += K2_THREADS) {
                    crops32[(size_t)cell * 4096 + i] = 0.f;
                    if (crops64) crops64[(size_t)cell * 4096 + i] = 0.0;
                }
            }
            continue;
        }
        my_cls = need <= lo_bytes ? 0 : (need <= hi_bytes ? 1 : 2);
        if (my_cls != cls) continue;

        unsigned char* base = cls == 2 ? gscratch + (size_t)blockIdx.x * gscratch_per_cta : dyn;
        const int hw = g.h * g.w;
        uint8_t* bins = base;
        uint16_t* rimg = reinterpret_cast<uint16_t*>(base + align16(hw));
        uint16_t* maps = reinterpret_cast<uint16_t*>(base + align16(hw) + align16(2 * (size_t)hw));
        double* T = reinterpret_cast<double*>(maps);

        const uint16_t* img = images + ((size_t)C.field * H + C.minr) * (size_t)W + C.minc;

        // ---- A: load bbox, min / max ----
        int mn = 65535, mx = 0;
        for (int i = tid; i < hw; i += K2_THREADS) {
            const int y = i / g.w, x = i - y * g.w;
            const int v = __ldg(img + (size_t)y * W + x);
            rimg[i] = (uint16_t)v;
            mn = min(mn, v); mx = max(mx, v);
        }
        block_minmax(mn, mx, red_s);

        // ---- B: 14-bit quantise (round half even) and bin ----
        {
            const double inv = 1.0 / 65535.0;
            const double vmin = __dmul_rn((double)mn, inv), vmax = __dmul_rn((double)mx, inv);
            const double den = __dsub_rn(vmax, vmin);
            for (int i = tid; i < hw; i += K2_THREADS) {
                const double v = __dmul_rn((double)rimg[i], inv);
                int q;
                if (mn != mx) {
                    const double t = __dmul_rn(__ddiv_rn(__dsub_rn(v, vmin), den), 16383.0);
                    q = __double2int_rn(t);
                } else {
                    q = __double2int_rn(fmin(fmax(v, 0.0), 16383.0));
                }
                bins[i] = (uint8_t)(q / BIN_SIZE);
            }
        }
        __syncthreads();

        // ---- C: per-tile histogram -> clipped mapping ----
        const int ntiles = g.ntr * g.ntc;
        for (int t = wid; t < ntiles; t += K2_WARPS) {
            const int ti = t / g.ntc, tj = t - ti * g.ntc;
            uint32_t* hs = hist_s[wid];
#pragma unroll
            for (int m = 0; m < 8; ++m) hs[lane + 32 * m] = 0;
            __syncwarp();
            for (int p = lane; p < g.npix; p += 32) {
                const int a = p / g.kw, b = p - a * g.kw;
                const int y = reflect_idx(ti * g.kh + a, g.h);
                const int x = reflect_idx(tj * g.kw + b, g.w);
                atomicAdd(&hs[bins[y * g.w + x]], 1u);
            }
            __syncwarp();
            int hv[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) hv[m] = (int)hs[lane + 32 * m];
            __syncwarp();
            clip_and_map(hv, g.clim, g.npix, lane, maps + (size_t)t * NBINS);
        }
        __syncthreads();

        // ---- D: 4-corner interpolation, fp32 accumulation, truncate to uint16 ----
        int rmn = 65535, rmx = 0;
        {
            const int pr = g.kh / 2, pc = g.kw / 2;
            for (int i = tid; i < hw; i += K2_THREADS) {
                const int y = i / g.w, x = i - y * g.w;
                const int Y = y + pr, X = x + pc;
                const int I = Y / g.kh, a = Y - I * g.kh;
                const int J = X / g.kw, b = X - J * g.kw;
                const double cr = __ddiv_rn((double)a, (double)g.kh);
                const double cc = __ddiv_rn((double)b, (double)g.kw);
                const double icr = __dsub_rn(1.0, cr), icc = __dsub_rn(1.0, cc);
                const int t0 = min(max(I - 1, 0), g.ntr - 1), t1 = min(max(I, 0), g.ntr - 1);
                const int u0 = min(max(J - 1, 0), g.ntc - 1), u1 = min(max(J, 0), g.ntc - 1);
                const int bin = bins[i];
                const double m00 = (double)maps[((size_t)t0 * g.ntc + u0) * NBINS + bin];
                const double m01 = (double)maps[((size_t)t0 * g.ntc + u1) * NBINS + bin];
                const double m10 = (double)maps[((size_t)t1 * g.ntc + u0) * NBINS + bin];
                const double m11 = (double)maps[((size_t)t1 * g.ntc + u1) * NBINS + bin];
                float res = 0.f;
                res = __fadd_rn(res, __double2float_rn(__dmul_rn(m00, __dmul_rn(icc, icr))));
                res = __fadd_rn(res, __double2float_rn(__dmul_rn(m01, __dmul_rn(cc, icr))));
                res = __fadd_rn(res, __double2float_rn(__dmul_rn(m10, __dmul_rn(icc, cr))));
                res = __fadd_rn(res, __double2float_rn(__dmul_rn(m11, __dmul_rn(cc, cr))));
                const int r = (int)res;
                rimg[i] = (uint16_t)r;
                rmn = min(rmn, r); rmx = max(rmx, r);
            }
        }
        block_minmax(rmn, rmx, red_s);   // also orders the rimg writes / maps reads
        if (levels_out) {                // test tap: the bit-exact integer core
            uint16_t* dst = levels_out + level_offsets[cell];
            for (int i = tid; i < hw; i += K2_THREADS) dst[i] = rimg[i];
        }

        // ---- E: gaussian weights for the axes that shrink ----
        const double fr = __ddiv_rn((double)g.h, 64.0), fc = __ddiv_rn((double)g.w, 64.0);
        double sig_r = __dmul_rn(__dsub_rn(fr, 1.0), 0.5), sig_c = __dmul_rn(__dsub_rn(fc, 1.0), 0.5);
        if (sig_r < 0.0) sig_r = 0.0;
        if (sig_c < 0.0) sig_c = 0.0;
        const bool blur_r = sig_r > 1e-15, blur_c = sig_c > 1e-15;
        const int rad_r = blur_r ? (int)(4.0 * sig_r + 0.5) : 0;
        const int rad_c = blur_c ? (int)(4.0 * sig_c + 0.5) : 0;
        if (wid < 2) {
            const bool on = wid == 0 ? blur_r : blur_c;
            const int rad = wid == 0 ? rad_r : rad_c;
            const double sg = wid == 0 ? sig_r : sig_c;
            if (on) {
                double s = 0.0;
                const double cf = -0.5 / (sg * sg);
                for (int j = lane; j <= 2 * rad; j += 32) {
                    const double xx = (double)(j - rad);
                    const double e = exp(cf * xx * xx);
                    gw_s[wid][j] = e;
                    s += e;
                }
                s = warp_sum(s);
                __syncwarp();
                for (int j = lane; j <= 2 * rad; j += 32) gw_s[wid][j] = gw_s[wid][j] / s;
            }
        }
        __syncthreads();

        const bool degenerate = rmn == rmx;
        const double den = (double)(rmx - rmn);
        const double lo = degenerate ? fmin(fmax((double)rmn, 0.0), 1.0) : 0.0;
        const double hi = degenerate ? lo : 1.0;
        auto value = [&](int y, int x) -> double {
            const double r = (double)rimg[y * g.w + x];
            return degenerate ? fmin(fmax(r, 0.0), 1.0) : __ddiv_rn(r - (double)rmn, den);
        };

        // ---- F: rows: (zoom o gaussian) along axis 0 -> T[64][w] ----
        for (int i = tid; i < CIA_CROP * g.w; i += K2_THREADS) {
            const int oy = i / g.w, x = i - oy * g.w;
            const double cc = ((double)oy + 0.5) * fr - 0.5;
            const double fl = floor(cc);
            const double t = cc - fl;
            const int i0 = (int)fl;
            double v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int yy = i0 + e;
                if (!blur_r) {
                    v[e] = value(mirror_idx(yy, g.h), x);
                } else {
                    double acc = 0.0;
                    for (int j = 0; j <= 2 * rad_r; ++j)
                        acc += gw_s[0][j] * value(mirror_idx(mirror_idx(yy, g.h) + j - rad_r, g.h), x);
                    v[e] = acc;
                }
            }
            T[i] = (1.0 - t) * v[0] + t * v[1];
        }
        __syncthreads();

        // ---- G: columns -> 64x64, clip, store ----
        for (int i = tid; i < CIA_CROP * CIA_CROP; i += K2_THREADS) {
            const int oy = i >> 6, ox = i & 63;
            const double cc = ((double)ox + 0.5) * fc - 0.5;
            const double fl = floor(cc);
            const double t = cc - fl;
            const int j0 = (int)fl;
            const double* Trow = T + (size_t)oy * g.w;
            double v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int xx = j0 + e;
                if (!blur_c) {
                    v[e] = Trow[mirror_idx(xx, g.w)];
                } else {
                    double acc = 0.0;
                    for (int j = 0; j <= 2 * rad_c; ++j)
                        acc += gw_s[1][j] * Trow[mirror_idx(mirror_idx(xx, g.w) + j - rad_c, g.w)];
                    v[e] = acc;
                }
            }
            double o = (1.0 - t) * v[0] + t * v[1];
            o = fmin(fmax(o, lo), hi);
            crops32[(size_t)cell * 4096 + i] = (float)o;          // det:122 rounding
            if (crops64) crops64[(size_t)cell * 4096 + i] = o;
        }
        __syncthreads();
    }
}

}  // namespace

int k_crop_resize(cia_ctx* h, const uint16_t* images, int H, int W, const cia_cell* cells,
                  int n_cells, const int32_t* n_cells_dev, const cia_params* p, float* crops32,
                  double* crops64, cudaStream_t s, uint16_t* levels_out,
                  const int64_t* level_offsets) {
    if (n_cells <= 0) return CIA_OK;
    static bool attr_set = false;
    const size_t lo_bytes = 94 * 1024, hi_bytes = 208 * 1024;   // 2 CTAs/SM and 1 CTA/SM with the 17 KB static part
    if (!attr_set) {
        CIA_CUDA(cudaFuncSetAttribute(crop_clahe_resize_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hi_bytes));
        attr_set = true;
    }
    const int huge_ctas = 32;
    const size_t per_cta = cell_bytes(MAX_SIDE, MAX_SIDE, 15, 15);
    int rc = ws_reserve(h, h->ws_crop_scratch, per_cta * huge_ctas);
    if (rc) return rc;
    int g0 = h->num_sms * 2; if (g0 > n_cells) g0 = n_cells;
    int g1 = h->num_sms;     if (g1 > n_cells) g1 = n_cells;
    int g2 = huge_ctas;      if (g2 > n_cells) g2 = n_cells;
    crop_clahe_resize_kernel<<<g0, K2_THREADS, lo_bytes, s>>>(
        images, H, W, cells, n_cells, n_cells_dev, p->clip_limit, crops32, crops64, 0, lo_bytes,
        hi_bytes, nullptr, 0, h->status_dev, levels_out, level_offsets);
    CIA_LAUNCH_CHECK();
    crop_clahe_resize_kernel<<<g1, K2_THREADS, hi_bytes, s>>>(
        images, H, W, cells, n_cells, n_cells_dev, p->clip_limit, crops32, crops64, 1, lo_bytes,
        hi_bytes, nullptr, 0, h->status_dev, levels_out, level_offsets);
    CIA_LAUNCH_CHECK();
    crop_clahe_resize_kernel<<<g2, K2_THREADS, 0, s>>>(
        images, H, W, cells, n_cells, n_cells_dev, p->clip_limit, crops32, crops64, 2, lo_bytes,
        hi_bytes, (unsigned char*)h->ws_crop_scratch.p, per_cta, h->status_dev, levels_out, level_offsets);
    CIA_LAUNCH_CHECK();
    return CIA_OK;
}
