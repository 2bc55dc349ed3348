// scan.cu -- K1 label scan (regionprops bbox / area / raw moments) and K1b quality
// gates + ordered compaction.
//
// Replaces improved_detection.py:67 (regionprops), :72-95 (border, area, eccentricity,
// mean/std gates) -- training twin CAE_improved_modeltrain.py:59-88.
//
// K1 is HBM-bound integer work: every label pixel is read once with 16-byte loads
// (4 B / pixel algorithmic traffic); a warp owns a 128-pixel row segment, groups its
// lanes by label with match.any and issues one set of atomics per (segment, label).
#include "common.cuh"

#include <climits>

namespace {

constexpr int SCAN_THREADS = 256;

// During accumulation the bbox fields hold complements so that a zeroed table is the
// identity: minr <- max(H - r), minc <- max(W - c), maxr <- max(r + 1), maxc <- max(c + 1).
__global__ void __launch_bounds__(SCAN_THREADS)
label_scan_kernel(const int32_t* __restrict__ labels, int n_fields, int H, int W, int max_label,
                  cia_region* __restrict__ regions, int32_t* status, int vec_ok) {
    const int lane = threadIdx.x & 31;
    const int warps_per_block = SCAN_THREADS / 32;
    const int segs = (W + 127) >> 7;
    const int f = blockIdx.y;                         // one field per grid row: no 64-bit divisions
    const int row_stride = gridDim.x * warps_per_block;
    cia_region* tab = regions + (size_t)f * max_label;

    for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < H; r += row_stride) {
      const int32_t* row = labels + ((size_t)f * H + r) * (size_t)W;
      // software pipeline: the next segment's 16-byte load is in flight while this one is reduced
      int4 nxt = make_int4(0, 0, 0, 0);
      if (vec_ok && (lane << 2) + 3 < W) nxt = __ldg(reinterpret_cast<const int4*>(row + (lane << 2)));
      for (int seg = 0; seg < segs; ++seg) {
        const int c0 = (seg << 7) + (lane << 2);
        int lab[4];
        if (vec_ok) {
            lab[0] = nxt.x; lab[1] = nxt.y; lab[2] = nxt.z; lab[3] = nxt.w;
            const int cn = c0 + 128;
            nxt = make_int4(0, 0, 0, 0);
            if (seg + 1 < segs && cn + 3 < W) nxt = __ldg(reinterpret_cast<const int4*>(row + cn));
            if (c0 + 3 >= W) { lab[0] = lab[1] = lab[2] = lab[3] = 0; }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) lab[k] = (c0 + k < W) ? __ldg(row + c0 + k) : 0;
        }
        // background-only segments (the majority) leave before the range check: 0 is in range
        const unsigned any = __ballot_sync(0xffffffffu, (lab[0] | lab[1] | lab[2] | lab[3]) != 0);
        if (any == 0) continue;
        // negative labels are background, as scipy.ndimage.find_objects (behind regionprops, det:67)
        // treats them; a label above max_label cannot come from labels.max() and is reported
        bool bad = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (lab[k] > max_label) bad = true;
            if ((unsigned)lab[k] > (unsigned)max_label) lab[k] = 0;
        }
        if (bad) raise_status(status, CIA_E_LABEL);

#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool start = lab[j] != 0 && (j == 0 || lab[j] != lab[j - 1]);
            const unsigned act = __ballot_sync(0xffffffffu, start);
            if (act == 0) continue;
            if (start) {
                const int l = lab[j];
                uint32_t cnt = 0, sc = 0, sc2 = 0;
                // contiguous run [j, e) of this lane's four pixels
                int e = j + 1;
#pragma unroll
                for (int k = 1; k < 4; ++k)
                    if (k > j && e == k && lab[k] == l) e = k + 1;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (k >= j && k < e) {
                        const uint32_t c = (uint32_t)(c0 + k);
                        cnt += 1; sc += c; sc2 += c * c;
                    }
                }
                const int clast = c0 + e - 1;
                const unsigned m = __match_any_sync(act, l);
                const int leader = __ffs(m) - 1;
                const uint32_t tcnt = __reduce_add_sync(m, cnt);
                const uint32_t tsc = __reduce_add_sync(m, sc);
                const uint32_t lo = __reduce_add_sync(m, sc2 & 0xFFFFu);
                const uint32_t hi = __reduce_add_sync(m, sc2 >> 16);
                const int cmax1 = __reduce_max_sync(m, clast + 1);
                const int cminc = __reduce_max_sync(m, W - (c0 + j));
                if (lane == leader) {
                    cia_region* R = tab + (l - 1);
                    const unsigned long long rr = (unsigned long long)r;
                    atomicAdd(&R->area, tcnt);
                    atomicMax(&R->minr, H - r);
                    atomicMax(&R->maxr, r + 1);
                    atomicMax(&R->minc, cminc);
                    atomicMax(&R->maxc, cmax1);
                    atomicAdd((unsigned long long*)&R->m10, rr * tcnt);
                    atomicAdd((unsigned long long*)&R->m01, (unsigned long long)tsc);
                    atomicAdd((unsigned long long*)&R->m20, rr * rr * tcnt);
                    atomicAdd((unsigned long long*)&R->m02,
                              (unsigned long long)lo + ((unsigned long long)hi << 16));
                    atomicAdd((unsigned long long*)&R->m11, rr * tsc);
                }
            }
        }
      }
    }
}

// K1 on run-length encoded label fields (the transport format of transport.cu): regionprops'
// bbox / area / raw moments are sums over pixels, and over a run (row r, columns [x0, x1), one
// label) they have closed forms, so the region table comes straight from the runs -- the dense
// int32 field (16.8 MB per 2048^2 field) is neither rebuilt in HBM nor read back.  One thread per
// run (grid-stride over the field's run list), row found by binary search in row_off.  Same
// table, same complement convention and same label check as label_scan_kernel: bit-identical.
__global__ void __launch_bounds__(SCAN_THREADS)
label_scan_rle_kernel(const uint32_t* __restrict__ slots, size_t slot_words, int H, int W, int max_label,
                      cia_region* __restrict__ regions, int32_t* status) {
    const int f = blockIdx.y;
    const uint32_t* slot = slots + (size_t)f * slot_words;
    const uint32_t* runs = slot + ((H + 2) & ~1);     // (x0, label) word pairs; 4-byte loads: a caller's
                                                      // slot stride may be odd
    const uint32_t n_runs = __ldg(slot + H);
    const size_t cap_runs = (slot_words - (size_t)((H + 2) & ~1)) / 2;
    if (n_runs > cap_runs) { if (threadIdx.x == 0 && blockIdx.x == 0) raise_status(status, CIA_E_ARG); return; }
    cia_region* tab = regions + (size_t)f * max_label;
    for (uint32_t j = blockIdx.x * SCAN_THREADS + threadIdx.x; j < n_runs; j += gridDim.x * SCAN_THREADS) {
        const int l = (int)__ldg(runs + 2 * (size_t)j + 1);
        if (l <= 0) continue;                                 // negative labels: background (find_objects)
        if (l > max_label) { raise_status(status, CIA_E_LABEL); continue; }
        int lo = 0, hi = H;                                   // largest r with row_off[r] <= j
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(slot + mid) <= j) lo = mid; else hi = mid;
        }
        const int r = lo;
        const int x0 = (int)__ldg(runs + 2 * (size_t)j);
        int x1 = W;
        if (j + 1 < __ldg(slot + r + 1)) x1 = min(W, (int)__ldg(runs + 2 * (size_t)j + 2));
        if (x0 < 0 || x1 <= x0) continue;
        const long long a = x0, b = x1;
        const unsigned long long cnt = (unsigned long long)(b - a);
        const unsigned long long sc = (unsigned long long)((a + b - 1) * (b - a) / 2);
        const unsigned long long sc2 = (unsigned long long)(((b - 1) * b * (2 * b - 1) - (a - 1) * a * (2 * a - 1)) / 6);
        const unsigned long long rr = (unsigned long long)r;
        cia_region* R = tab + (l - 1);
        atomicAdd(&R->area, (uint32_t)cnt);
        atomicMax(&R->minr, H - r);
        atomicMax(&R->maxr, r + 1);
        atomicMax(&R->minc, W - x0);
        atomicMax(&R->maxc, x1);
        atomicAdd((unsigned long long*)&R->m10, rr * cnt);
        atomicAdd((unsigned long long*)&R->m01, sc);
        atomicAdd((unsigned long long*)&R->m20, rr * rr * cnt);
        atomicAdd((unsigned long long*)&R->m02, sc2);
        atomicAdd((unsigned long long*)&R->m11, rr * sc);
    }
}

__global__ void finalize_regions_kernel(cia_region* regions, long long n_slots, int H, int W) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots) return;
    cia_region* R = regions + i;
    if (R->area == 0) return;
    R->minr = H - R->minr;
    R->minc = W - R->minc;
}

// ---- K1b: gates -----------------------------------------------------------
// One warp per region slot.  Side output stats[slot] = {eccentricity, mean, std}.
__global__ void __launch_bounds__(256)
gate_kernel(const uint16_t* __restrict__ images, int n_fields, int H, int W, int max_label,
            cia_region* __restrict__ regions, cia_params p, double* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const long long n_slots = (long long)n_fields * max_label;
    long long slot = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long stride = ((long long)gridDim.x * blockDim.x) >> 5;
    for (; slot < n_slots; slot += stride) {
        cia_region* R = regions + slot;
        const uint32_t area = R->area;
        if (area == 0) continue;
        const int minr = R->minr, minc = R->minc, maxr = R->maxr, maxc = R->maxc;
        bool keep = !(minr < p.border_margin || minc < p.border_margin ||
                      maxr > H - p.border_margin || maxc > W - p.border_margin);      // det:76
        keep = keep && !((int)area < p.area_min || (int)area > p.area_max);           // det:80
        double ecc = 0.0;
        if (keep) {
            // exact integer n*mu_pq, then the closed-form eigenvalues of the inertia tensor
            const long long n = (long long)area;
            const long long m10 = (long long)R->m10, m01 = (long long)R->m01;
            const long long a20 = n * (long long)R->m20 - m10 * m10;
            const long long a02 = n * (long long)R->m02 - m01 * m01;
            const long long a11 = n * (long long)R->m11 - m10 * m01;
            const double nn = __dmul_rn((double)n, (double)n);
            const double a = __ddiv_rn((double)a02, nn);
            const double c = __ddiv_rn((double)a20, nn);
            const double b = -__ddiv_rn((double)a11, nn);
            const double half_tr = __dmul_rn(0.5, __dadd_rn(a, c));
            const double hd = __dmul_rn(0.5, __dsub_rn(a, c));
            const double rad = __dsqrt_rn(__dadd_rn(__dmul_rn(hd, hd), __dmul_rn(b, b)));
            const double l1 = __dadd_rn(half_tr, rad);
            double l2 = __dsub_rn(half_tr, rad);
            if (l2 < 0.0) l2 = 0.0;
            if (l1 > 0.0) {
                double t = __dsub_rn(1.0, __ddiv_rn(l2, l1));
                if (t < 0.0) t = 0.0;
                ecc = __dsqrt_rn(t);
            }
            keep = !(ecc > p.ecc_max);                                                // det:84
        }
        double mean = 0.0, sd = 0.0;
        if (keep) {
            const int f = (int)(slot / max_label);
            const int h = maxr - minr, w = maxc - minc;
            const uint16_t* img = images + ((size_t)f * H + minr) * (size_t)W + minc;
            const int npx = h * w;
            unsigned long long s = 0;
            for (int i = lane; i < npx; i += 32) {
                const int y = i / w, x = i - y * w;
                s += __ldg(img + (size_t)y * W + x);
            }
            s = warp_sum(s);
            mean = __ddiv_rn((double)s, (double)npx);                                 // det:91
            double q = 0.0;
            for (int i = lane; i < npx; i += 32) {
                const int y = i / w, x = i - y * w;
                const double d = __dsub_rn((double)__ldg(img + (size_t)y * W + x), mean);
                q = __dadd_rn(q, __dmul_rn(d, d));
            }
            q = warp_sum(q);
            sd = __dsqrt_rn(__ddiv_rn(q, (double)npx));                               // det:92
            keep = !(mean < p.mean_min || sd < p.std_min);                            // det:94
        }
        if (lane == 0) {
            R->flags = keep ? 1 : 0;
            if (keep) {
                stats[slot * 3 + 0] = ecc;
                stats[slot * 3 + 1] = mean;
                stats[slot * 3 + 2] = sd;
            }
        }
    }
}

// ---- ordered compaction -----------------------------------------------------
__global__ void __launch_bounds__(256)
count_kernel(const cia_region* __restrict__ regions, int max_label, int32_t* __restrict__ counts) {
    const int f = blockIdx.x;
    const cia_region* tab = regions + (size_t)f * max_label;
    int c = 0;
    for (int i = threadIdx.x; i < max_label; i += blockDim.x) c += (tab[i].area != 0) & tab[i].flags;
    __shared__ int sh[8];
    c = warp_sum(c);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < 8; ++i) t += sh[i];
        counts[f] = t;
    }
}

__global__ void __launch_bounds__(256)
scatter_kernel(const cia_region* __restrict__ regions, const double* __restrict__ stats,
               int n_fields, int max_label, const int32_t* __restrict__ counts,
               cia_cell* __restrict__ cells, int cells_cap, int32_t* n_cells_dev,
               int32_t* field_counts_out, int32_t* status) {
    const int f = blockIdx.x;
    __shared__ int sh[8];
    __shared__ int base_sh;
    // offset of this field = sum of the counts of earlier fields
    int part = 0;
    for (int i = threadIdx.x; i < f; i += blockDim.x) part += counts[i];
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < 8; ++i) t += sh[i];
        base_sh = t;
        if (field_counts_out) field_counts_out[f] = counts[f];
        if (f == n_fields - 1) {
            const int total = t + counts[f];
            *n_cells_dev = total;
            if (total > cells_cap) raise_status(status, CIA_E_CAPACITY);
        }
    }
    __syncthreads();
    int base = base_sh;
    const cia_region* tab = regions + (size_t)f * max_label;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i0 = 0; i0 < max_label; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        const bool keep = i < max_label && tab[i].area != 0 && (tab[i].flags & 1);
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        __syncthreads();
        if (lane == 0) sh[wid] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
        for (int k = 0; k < 8; ++k) { if (k < wid) woff += sh[k]; tot += sh[k]; }
        if (keep) {
            const int pos = base + woff + __popc(bal & ((1u << lane) - 1u));
            if (pos < cells_cap) {
                const cia_region R = tab[i];
                cia_cell c;
                c.field = f; c.label = i + 1;
                c.minr = R.minr; c.minc = R.minc; c.maxr = R.maxr; c.maxc = R.maxc;
                c.area = (int32_t)R.area; c.pad_ = 0;
                const double* st = stats + ((size_t)f * max_label + i) * 3;
                c.eccentricity = st[0]; c.mean_intensity = st[1]; c.std_intensity = st[2];
                cells[pos] = c;
            }
        }
        base += tot;
    }
}

}  // namespace

// ---- K1c: solidity (det:106, train:101) --------------------------------------
// skimage regionprops: solidity = area / area_convex, area_convex = convex_hull_image(mask).sum() with
// offset_coordinates (every mask pixel contributes the four midpoints of its edges) and
// include_borders (pixel centres ON the hull count).  The hull of a pixel set depends only on the
// leftmost / rightmost pixel of every row, so: lanes find the row extents, doubled coordinates keep
// everything integer (edge midpoints (2r +- 1, 2c), (2r, 2c +- 1); pixel centres (2r, 2c)), one lane
// runs the two monotone chains (right = upper envelope of X over Y, left = lower envelope) and
// counts per pixel row the even X between the envelopes with exact rational floor / ceil.
// One warp per cell; scratch per warp: 6 * (2h + 1) ints in global memory.
__global__ void __launch_bounds__(128)
solidity_kernel(const int32_t* __restrict__ labels, int H, int W, const cia_cell* __restrict__ cells, int n_cells,
                const int32_t* __restrict__ n_dev, double* __restrict__ out, int32_t* __restrict__ scratch,
                size_t scratch_per_warp) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int n = dev_count(n_cells, n_dev);
    int32_t* base = scratch + (size_t)warp * scratch_per_warp;
    for (int cell = warp; cell < n; cell += nwarps) {
        const cia_cell C = cells[cell];
        const int h = C.maxr - C.minr, w = C.maxc - C.minc;
        const int K = 2 * h + 1;                           // Y = 2r - 1 .. 2r + 1 over the bbox rows, index k = Y + 1
        int32_t* xr = base;            int32_t* xl = base + K;
        int32_t* hr = base + 2 * K;    int32_t* hl = base + 4 * K;      // hull stacks: (k, x) pairs
        const int32_t* lab = labels + ((size_t)C.field * H + C.minr) * (size_t)W + C.minc;
        for (int k = lane; k < K; k += 32) { xr[k] = INT_MIN; xl[k] = INT_MAX; }
        __syncwarp();
        // row extents; columns are local + 1 so that doubled coordinates stay positive
        for (int r = lane; r < h; r += 32) {
            int cl = -1, cr = -1;
            const int32_t* row = lab + (size_t)r * W;
            for (int c = 0; c < w; ++c)
                if (__ldg(row + c) == C.label) { if (cl < 0) cl = c; cr = c; }
            if (cl >= 0) {
                const int L = 2 * (cl + 1), Rr = 2 * (cr + 1);
                atomicMax(&xr[2 * r + 1], Rr + 1); atomicMin(&xl[2 * r + 1], L - 1);      // (2r, 2c +- 1)
                atomicMax(&xr[2 * r], Rr);     atomicMin(&xl[2 * r], L);                 // (2r - 1, 2c)
                atomicMax(&xr[2 * r + 2], Rr); atomicMin(&xl[2 * r + 2], L);             // (2r + 1, 2c)
            }
        }
        __syncwarp();
        if (lane == 0) {
            // monotone chains over k ascending: keep the envelope convex (cross product test, 64-bit)
            int nr = 0, nl = 0;
            for (int k = 0; k < K; ++k) {
                if (xr[k] == INT_MIN) continue;
                const long long px = xr[k];
                while (nr >= 2) {
                    const long long ax = hr[2 * (nr - 2) + 1], ak = hr[2 * (nr - 2)], bx = hr[2 * (nr - 1) + 1], bk = hr[2 * (nr - 1)];
                    // b lies on or below the chord a -> p: not a vertex of the upper envelope
                    if ((bx - ax) * (k - ak) <= (px - ax) * (bk - ak)) --nr; else break;
                }
                hr[2 * nr] = k; hr[2 * nr + 1] = (int)px; ++nr;
                const long long qx = xl[k];
                while (nl >= 2) {
                    const long long ax = hl[2 * (nl - 2) + 1], ak = hl[2 * (nl - 2)], bx = hl[2 * (nl - 1) + 1], bk = hl[2 * (nl - 1)];
                    if ((bx - ax) * (k - ak) >= (qx - ax) * (bk - ak)) --nl; else break;
                }
                hl[2 * nl] = k; hl[2 * nl + 1] = (int)qx; ++nl;
            }
            long long area_convex = 0;
            int ir = 0, il = 0;
            for (int r = 0; r < h && nr > 0; ++r) {
                const int k = 2 * r + 1;                   // Y = 2r
                if (k < hr[0] || k > hr[2 * (nr - 1)]) continue;
                while (ir + 1 < nr - 1 && hr[2 * (ir + 1)] <= k) ++ir;
                while (il + 1 < nl - 1 && hl[2 * (il + 1)] <= k) ++il;
                long long xmax2, xmin2;                    // floor(X_right / 2), ceil(X_left / 2) in doubled local columns
                if (nr == 1) { xmax2 = hr[1] / 2; xmin2 = (hl[1] + 1) / 2; }
                else {
                    const long long k0 = hr[2 * ir], x0 = hr[2 * ir + 1], k1 = hr[2 * ir + 2], x1 = hr[2 * ir + 3], d = k1 - k0;
                    xmax2 = (x0 * d + (x1 - x0) * (k - k0)) / (2 * d);                     // numerator >= 0
                    const long long m0 = hl[2 * il], y0 = hl[2 * il + 1], m1 = hl[2 * il + 2], y1 = hl[2 * il + 3], e = m1 - m0;
                    const long long num = y0 * e + (y1 - y0) * (k - m0);
                    xmin2 = (num + 2 * e - 1) / (2 * e);
                }
                if (xmax2 >= xmin2) area_convex += xmax2 - xmin2 + 1;
            }
            out[cell] = area_convex > 0 ? (double)C.area / (double)area_convex : 0.0;
        }
        __syncwarp();
    }
}

int k_solidity(cia_ctx* h, const int32_t* labels, int H, int W, const cia_cell* cells, int n_cells,
               const int32_t* n_dev, double* out, cudaStream_t s) {
    if (n_cells <= 0) return CIA_OK;
    if (H <= 0 || W <= 0) { h->err = "cia_solidity: bad shape"; return CIA_E_ARG; }
    int blocks = (n_cells + 3) / 4;
    if (blocks > h->num_sms * 4) blocks = h->num_sms * 4;
    const size_t per_warp = 6 * (size_t)(2 * H + 1);
    int rc = ws_reserve(h, h->ws_flags, (size_t)blocks * 4 * per_warp * sizeof(int32_t));
    if (rc) return rc;
    solidity_kernel<<<blocks, 128, 0, s>>>(labels, H, W, cells, n_cells, n_dev, out, (int32_t*)h->ws_flags.p, per_warp);
    CIA_LAUNCH_CHECK();
    return CIA_OK;
}

int k_label_scan(cia_ctx* h, const int32_t* labels, int n_fields, int H, int W, int max_label,
                 cia_region* regions, cudaStream_t s) {
    if (n_fields <= 0 || H <= 0 || W <= 0 || max_label <= 0 || H > 32767 || W > 32767) {
        h->err = "cia_label_scan: bad shape (need 0 < H,W <= 32767, max_label > 0)";
        return CIA_E_ARG;
    }
    const size_t n_slots = (size_t)n_fields * max_label;
    CIA_CUDA(cudaMemsetAsync(regions, 0, n_slots * sizeof(cia_region), s));
    const int vec_ok = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(labels) & 15) == 0);
    if (n_fields > 65535) { h->err = "cia_label_scan: at most 65535 fields per call"; return CIA_E_ARG; }
    int bx = (H + 7) / 8;
    const int want = (h->num_sms * 16 + n_fields - 1) / n_fields;      // ~16 resident blocks per SM overall
    if (bx > want) bx = want;
    if (bx < 1) bx = 1;
    label_scan_kernel<<<dim3(bx, n_fields), SCAN_THREADS, 0, s>>>(labels, n_fields, H, W, max_label,
                                                                 regions, h->status_dev, vec_ok);
    CIA_LAUNCH_CHECK();
    finalize_regions_kernel<<<(int)((n_slots + 255) / 256), 256, 0, s>>>(regions, (long long)n_slots, H, W);
    CIA_LAUNCH_CHECK();
    return CIA_OK;
}

int k_label_scan_rle(cia_ctx* h, const uint32_t* slots, size_t slot_words, int n_fields, int H, int W,
                     int max_label, cia_region* regions, cudaStream_t s) {
    if (n_fields <= 0 || H <= 0 || W <= 0 || max_label <= 0 || H > 32767 || W > 32767 ||
        slot_words < (size_t)((H + 2) & ~1) + 2) {
        h->err = "cia_label_scan_rle: bad shape (need 0 < H,W <= 32767, max_label > 0, slot_words >= H + 4)";
        return CIA_E_ARG;
    }
    if (n_fields > 65535) { h->err = "cia_label_scan_rle: at most 65535 fields per call"; return CIA_E_ARG; }
    const size_t n_slots = (size_t)n_fields * max_label;
    CIA_CUDA(cudaMemsetAsync(regions, 0, n_slots * sizeof(cia_region), s));
    int bx = (h->num_sms * 8 + n_fields - 1) / n_fields;
    if (bx > 128) bx = 128;
    if (bx < 1) bx = 1;
    label_scan_rle_kernel<<<dim3(bx, n_fields), SCAN_THREADS, 0, s>>>(slots, slot_words, H, W, max_label,
                                                                     regions, h->status_dev);
    CIA_LAUNCH_CHECK();
    finalize_regions_kernel<<<(int)((n_slots + 255) / 256), 256, 0, s>>>(regions, (long long)n_slots, H, W);
    CIA_LAUNCH_CHECK();
    return CIA_OK;
}

int k_filter(cia_ctx* h, const uint16_t* images, int n_fields, int H, int W, int max_label,
             cia_region* regions, const cia_params* p, cia_cell* cells, int cells_cap,
             int32_t* n_cells_dev, int32_t* field_counts_dev, cudaStream_t s) {
    if (n_fields <= 0 || max_label <= 0 || cells_cap < 0) {
        h->err = "cia_filter: bad argument";
        return CIA_E_ARG;
    }
    const size_t n_slots = (size_t)n_fields * max_label;
    int rc = ws_reserve(h, h->ws_flags, n_slots * 3 * sizeof(double) + (size_t)n_fields * sizeof(int32_t));
    if (rc) return rc;
    double* stats = (double*)h->ws_flags.p;
    int32_t* counts = (int32_t*)(stats + n_slots * 3);
    long long warps = (long long)n_slots;
    long long blocks = (warps + 7) / 8;
    const long long cap = (long long)h->num_sms * 16;
    if (blocks > cap) blocks = cap;
    gate_kernel<<<(int)blocks, 256, 0, s>>>(images, n_fields, H, W, max_label, regions, *p, stats);
    CIA_LAUNCH_CHECK();
    count_kernel<<<n_fields, 256, 0, s>>>(regions, max_label, counts);
    CIA_LAUNCH_CHECK();
    scatter_kernel<<<n_fields, 256, 0, s>>>(regions, stats, n_fields, max_label, counts, cells,
                                            cells_cap, n_cells_dev, field_counts_dev, h->status_dev);
    CIA_LAUNCH_CHECK();
    return CIA_OK;
}
