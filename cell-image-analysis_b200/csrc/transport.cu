// transport.cu -- run-length transport of label fields over PCIe.
//
// The reference hands StarDist's int32 label image to regionprops on the host
// (improved_detection.py:66-70).  A B200 screens a 2048 x 2048 field in ~0.3 ms, less than the
// 0.30 ms its 16.8 MB of int32 labels take over PCIe Gen5 -- the host->device copy, not a
// kernel, bounds the end-to-end rate.  Label images are piecewise constant along rows, so the
// host side run-length encodes them (multi-threaded, one pass at memory speed) and only the runs
// (~0.3 MB per field) cross the bus.  The fused path then builds the region table straight from
// the runs (label_scan_rle_kernel in scan.cu: labels are read by the scan only, and its sums have
// closed forms per run); rle_expand_kernel below rebuilds the dense int32 field in HBM for callers
// that want it.  Lossless: expand(encode(L)) == L bit for bit.
//
// Slot layout per field (uint32 words; slot stride chosen by the caller):
//   [0 .. H]      row_off: index of the first run of each row, row_off[H] = number of runs
//   [R0 ..]       runs as (x0, label) pairs, R0 = (H + 2) & ~1; a run extends to the next run's
//                 x0 or the end of the row; every row starts with a run at x0 = 0
//
// Patch transport of the IMAGE (round 2).  The device reads the uint16 image only inside the bounding boxes
// of labelled regions (gates: mean / std of the bbox, det:88-95; crop: det:88): ~1.2 MB of the 8.4 MB of a 2048^2
// field with ~500 cells.  The encoder thread that has just produced a field's runs therefore also packs the bbox
// rectangles of all present labels (bboxes from the runs, label ascending, rows contiguous) into a pinned
// staging slot; only that crosses PCIe, and patch_scatter_kernel writes the rectangles back into a dense
// [F, H, W] device buffer at their bbox positions, using the region table the device built from the same runs
// (same integers, same order).  Pixels outside every bbox are never read by the path and stay undefined.
// One GPU: the pass is no longer bound by the PCIe transfer of the images (8.9 -> 1.9 GB per 1024 fields).
#include "common.cuh"

#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>
#include <vector>

namespace {

inline size_t rle_runs_base(int H) { return (size_t)((H + 2) & ~1); }

}  // namespace

// host_rle.cpp (plain C++, AVX2 path selected at run time)
size_t cia_host_encode_field(const int32_t* lab, int H, int W, uint32_t* slot, size_t slot_words, int32_t* max_label);

namespace {

// One thread per four pixels: binary search for the run covering the first, then walk.
__global__ void __launch_bounds__(256)
rle_expand_kernel(const uint32_t* __restrict__ slots, size_t slot_words, int H, int W,
                  int32_t* __restrict__ labels) {
    const int f = blockIdx.z, y = blockIdx.y;
    const int x = 4 * (blockIdx.x * 256 + threadIdx.x);
    if (x >= W) return;
    const uint32_t* slot = slots + (size_t)f * slot_words;
    const uint32_t* runs = slot + ((H + 2) & ~1);               // (x0, label) word pairs; 4-byte loads only,
                                                                // a caller's slot stride may be odd
    uint32_t lo = __ldg(slot + y), hi = __ldg(slot + y + 1);    // runs of this row: [lo, hi)
    const uint32_t end = hi;
    while (hi - lo > 1) {                                       // largest j with runs[j].x0 <= x
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(runs + 2 * (size_t)mid) <= (uint32_t)x) lo = mid; else hi = mid;
    }
    uint32_t cur_label = __ldg(runs + 2 * (size_t)lo + 1);
    uint32_t next_x0 = lo + 1 < end ? __ldg(runs + 2 * (size_t)lo + 2) : 0xFFFFFFFFu;
    int32_t out[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        while ((uint32_t)(x + k) >= next_x0) {
            ++lo;
            cur_label = __ldg(runs + 2 * (size_t)lo + 1);
            next_x0 = lo + 1 < end ? __ldg(runs + 2 * (size_t)lo + 2) : 0xFFFFFFFFu;
        }
        out[k] = (int32_t)cur_label;
    }
    int32_t* dst = labels + ((size_t)f * H + y) * W + x;
    if (x + 4 <= W && (W & 3) == 0) {
        *reinterpret_cast<int4*>(dst) = make_int4(out[0], out[1], out[2], out[3]);
    } else {
        for (int k = 0; k < 4 && x + k < W; ++k) dst[k] = out[k];
    }
}


// One CTA per field: sizes of the present labels' bbox rectangles -> exclusive scan (= the packing order of
// cia_rle_encode_pack_fields) -> one warp per region copies its rows to the dense image.
__global__ void __launch_bounds__(256)
patch_scatter_kernel(const uint16_t* __restrict__ patches, size_t cap_px, const cia_region* __restrict__ regions,
                     int max_label, int H, int W, uint16_t* __restrict__ images, int32_t* status) {
    extern __shared__ uint32_t off_s[];              // [max_label + 1]
    __shared__ uint32_t part_s[256];
    const int f = blockIdx.x, tid = threadIdx.x;
    const cia_region* tab = regions + (size_t)f * max_label;
    const int chunk = (max_label + 255) / 256;
    const int l0 = tid * chunk, l1 = min(max_label, l0 + chunk);
    uint32_t sum = 0;
    for (int l = l0; l < l1; ++l) {
        const cia_region R = tab[l];
        const uint32_t sz = R.area ? (uint32_t)(R.maxr - R.minr) * (uint32_t)(R.maxc - R.minc) : 0u;
        off_s[l] = sum;                               // chunk-local exclusive prefix
        sum += sz;
    }
    part_s[tid] = sum;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1) {               // inclusive scan of the chunk sums
        const uint32_t v = tid >= d ? part_s[tid - d] : 0u;
        __syncthreads();
        part_s[tid] += v;
        __syncthreads();
    }
    const uint32_t base = tid ? part_s[tid - 1] : 0u;
    for (int l = l0; l < l1; ++l) off_s[l] += base;
    __syncthreads();
    if ((size_t)part_s[255] > cap_px) { if (tid == 0) raise_status(status, CIA_E_CAPACITY); return; }
    const uint16_t* src = patches + (size_t)f * cap_px;
    uint16_t* img = images + (size_t)f * H * W;
    const int lane = tid & 31;
    for (int l = tid >> 5; l < max_label; l += 8) {
        const cia_region R = tab[l];
        if (!R.area) continue;
        const int h = R.maxr - R.minr, w = R.maxc - R.minc;
        const uint16_t* ps = src + off_s[l];
        for (int y = 0; y < h; ++y) {
            uint16_t* row = img + (size_t)(R.minr + y) * W + R.minc;
            for (int x = lane; x < w; x += 32) row[x] = ps[y * w + x];
        }
    }
}

// Packs the bbox rectangles of one field (host).  Returns the pixels written, or SIZE_MAX if they do not fit.
size_t pack_field_patches(const uint16_t* img, int H, int W, const uint32_t* slot, int label_cap, uint16_t* out,
                          size_t cap_px, std::vector<int32_t>& box) {
    const uint32_t* runs = slot + rle_runs_base(H);
    box.assign((size_t)4 * (label_cap + 1), 0);       // minr, maxr (exclusive), minc, maxc (exclusive); maxr == 0: absent
    for (int l = 0; l <= label_cap; ++l) { box[4 * l] = H; box[4 * l + 2] = W; }
    for (int y = 0; y < H; ++y) {
        const uint32_t j0 = slot[y], j1 = slot[y + 1];
        for (uint32_t j = j0; j < j1; ++j) {
            const int32_t l = (int32_t)runs[2 * (size_t)j + 1];
            if (l <= 0 || l > label_cap) continue;    // background / out of range (the device reports the latter)
            const int x0 = (int)runs[2 * (size_t)j], x1 = j + 1 < j1 ? (int)runs[2 * (size_t)j + 2] : W;
            int32_t* b = &box[4 * (size_t)l];
            if (y < b[0]) b[0] = y;
            b[1] = y + 1;
            if (x0 < b[2]) b[2] = x0;
            if (x1 > b[3]) b[3] = x1;
        }
    }
    size_t n = 0;
    for (int l = 1; l <= label_cap; ++l) {
        const int32_t* b = &box[4 * (size_t)l];
        if (b[1] == 0) continue;
        const int h = b[1] - b[0], w = b[3] - b[2];
        if (n + (size_t)h * w > cap_px) return SIZE_MAX;
        for (int y = 0; y < h; ++y)
            std::memcpy(out + n + (size_t)y * w, img + (size_t)(b[0] + y) * W + b[2], (size_t)w * sizeof(uint16_t));
        n += (size_t)h * w;
    }
    return n;
}

}  // namespace

int k_patch_scatter(cia_ctx* h, const uint16_t* patches, size_t cap_px, const cia_region* regions, int n_fields,
                    int max_label, int H, int W, uint16_t* images, cudaStream_t s) {
    if (n_fields <= 0) return CIA_OK;
    const size_t smem = (size_t)(max_label + 1) * sizeof(uint32_t);
    if (smem > 200 * 1024) { h->err = "patch transport: max_label too large"; return CIA_E_UNSUPPORTED; }
    if (smem > 48 * 1024 && first_use(h, (const void*)patch_scatter_kernel))
        CIA_CUDA(cudaFuncSetAttribute(patch_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    patch_scatter_kernel<<<n_fields, 256, smem, s>>>(patches, cap_px, regions, max_label, H, W, images, h->status_dev);
    CIA_LAUNCH_CHECK();
    return CIA_OK;
}

namespace {
}  // namespace

extern "C" {

size_t cia_rle_slot_words(int H, int W) {
    // room for one run per 8 pixels on average (a quarter of the raw field); typical fields need ~3 %
    return rle_runs_base(H) + 2 * (((size_t)H * W + 7) / 8);
}

int cia_rle_encode_fields(const int32_t* labels_host, int n_fields, int H, int W, uint32_t* slots_host,
                          size_t slot_words, uint32_t* field_words, int32_t* max_label, int n_threads) {
    if (!labels_host || !slots_host || !field_words || n_fields < 0 || H <= 0 || W <= 0) return CIA_E_ARG;
    if (n_threads <= 0) {
        const char* e = getenv("CIA_HOST_THREADS");
        n_threads = e ? atoi(e) : (int)std::thread::hardware_concurrency();
        if (n_threads <= 0) n_threads = 1;
        if (n_threads > 32) n_threads = 32;
    }
    if (n_threads > n_fields) n_threads = n_fields > 0 ? n_fields : 1;
    std::atomic<int> next(0), overflow(0);
    std::vector<int32_t> mx((size_t)(n_fields > 0 ? n_fields : 1), 0);
    auto work = [&]() {
        for (;;) {
            const int f = next.fetch_add(1);
            if (f >= n_fields) break;
            const size_t w = cia_host_encode_field(labels_host + (size_t)f * H * W, H, W, slots_host + (size_t)f * slot_words,
                                          slot_words, &mx[f]);
            field_words[f] = (uint32_t)w;
            if (w == 0) overflow.store(1);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    if (max_label) {
        int32_t m = 0;
        for (int f = 0; f < n_fields; ++f) m = mx[f] > m ? mx[f] : m;
        *max_label = m;
    }
    return overflow.load() ? CIA_E_CAPACITY : CIA_OK;
}

// cia_rle_encode_fields + the patch transport of the image: the thread that encoded a field packs the bbox
// rectangles of its labels (1..label_cap) from images_host into patches_host[f * patch_cap_px ..]; patch_px[f] =
// pixels used, 0xFFFFFFFF if they do not fit (the caller then copies that chunk's images densely).
int cia_rle_encode_pack_fields(const int32_t* labels_host, const uint16_t* images_host, int n_fields, int H, int W,
                               uint32_t* slots_host, size_t slot_words, uint32_t* field_words, int32_t* max_label,
                               int label_cap, uint16_t* patches_host, size_t patch_cap_px, uint32_t* patch_px,
                               int n_threads) {
    if (!labels_host || !images_host || !slots_host || !field_words || !patches_host || !patch_px || n_fields < 0 ||
        H <= 0 || W <= 0 || label_cap <= 0) return CIA_E_ARG;
    if (n_threads <= 0) {
        const char* e = getenv("CIA_HOST_THREADS");
        n_threads = e ? atoi(e) : (int)std::thread::hardware_concurrency();
        if (n_threads <= 0) n_threads = 1;
        if (n_threads > 32) n_threads = 32;
    }
    if (n_threads > n_fields) n_threads = n_fields > 0 ? n_fields : 1;
    std::atomic<int> next(0), overflow(0);
    std::vector<int32_t> mx((size_t)(n_fields > 0 ? n_fields : 1), 0);
    auto work = [&]() {
        std::vector<int32_t> box;
        for (;;) {
            const int f = next.fetch_add(1);
            if (f >= n_fields) break;
            uint32_t* slot = slots_host + (size_t)f * slot_words;
            const size_t w = cia_host_encode_field(labels_host + (size_t)f * H * W, H, W, slot, slot_words, &mx[f]);
            field_words[f] = (uint32_t)w;
            patch_px[f] = 0xFFFFFFFFu;
            if (w == 0) { overflow.store(1); continue; }
            const size_t n = pack_field_patches(images_host + (size_t)f * H * W, H, W, slot, label_cap,
                                                patches_host + (size_t)f * patch_cap_px, patch_cap_px, box);
            if (n != SIZE_MAX) patch_px[f] = (uint32_t)n;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    if (max_label) {
        int32_t m = 0;
        for (int f = 0; f < n_fields; ++f) m = mx[f] > m ? mx[f] : m;
        *max_label = m;
    }
    return overflow.load() ? CIA_E_CAPACITY : CIA_OK;
}

// one async copy per field of exactly the pixels used
int cia_patch_upload(cia_handle h, const uint16_t* patches_host, int n_fields, size_t patch_cap_px,
                     const uint32_t* patch_px, uint16_t* patches_dev, void* stream) {
    if (!h) return CIA_E_ARG;
    if (!patches_host || !patch_px || !patches_dev) { h->err = "cia_patch_upload: null pointer"; return CIA_E_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    for (int f = 0; f < n_fields; ++f) {
        if (patch_px[f] == 0xFFFFFFFFu || patch_px[f] > patch_cap_px) { h->err = "cia_patch_upload: field not packed"; return CIA_E_ARG; }
        if (patch_px[f] == 0) continue;
        CIA_CUDA(cudaMemcpyAsync(patches_dev + (size_t)f * patch_cap_px, patches_host + (size_t)f * patch_cap_px,
                                 (size_t)patch_px[f] * sizeof(uint16_t), cudaMemcpyHostToDevice, s));
    }
    return CIA_OK;
}

// Streaming-read bandwidth of host memory as the run-length encoder sees it: n_threads threads each
// sum their contiguous share of `buf` with 8-byte loads, `reps` passes; returns GB/s (0 on bad input).
// bench.py runs it on the pinned label pool, alone and next to H2D copies, to say whether the
// end-to-end rate is bound by the host's cores or by its DRAM.
double cia_host_read_probe(const void* buf, size_t bytes, int n_threads, int reps) {
    if (!buf || bytes < 4096 || reps <= 0) return 0.0;
    if (n_threads <= 0) {
        const char* e = getenv("CIA_HOST_THREADS");
        n_threads = e ? atoi(e) : (int)std::thread::hardware_concurrency();
        if (n_threads <= 0) n_threads = 1;
        if (n_threads > 64) n_threads = 64;
    }
    const size_t words = bytes / 8, per = words / (size_t)n_threads;
    std::vector<uint64_t> sums((size_t)n_threads, 0);
    auto work = [&](int t) {
        const uint64_t* p = (const uint64_t*)buf + (size_t)t * per;
        uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        for (int r = 0; r < reps; ++r)
            for (size_t i = 0; i + 4 <= per; i += 4) { a0 += p[i]; a1 += p[i + 1]; a2 += p[i + 2]; a3 += p[i + 3]; }
        sums[(size_t)t] = a0 + a1 + a2 + a3;
    };
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& t : pool) t.join();
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    volatile uint64_t sink = 0;
    for (uint64_t v : sums) sink = sink + v;
    (void)sink;
    return dt > 0 ? (double)per * 8.0 * n_threads * reps / dt / 1e9 : 0.0;
}

int cia_rle_upload(cia_handle h, const uint32_t* slots_host, int n_fields, size_t slot_words,
                   const uint32_t* field_words, uint32_t* slots_dev, void* stream) {
    if (!h) return CIA_E_ARG;
    if (!slots_host || !field_words || !slots_dev) { h->err = "cia_rle_upload: null pointer"; return CIA_E_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    for (int f = 0; f < n_fields; ++f) {
        if (field_words[f] == 0 || field_words[f] > slot_words) { h->err = "cia_rle_upload: field not encoded"; return CIA_E_ARG; }
        CIA_CUDA(cudaMemcpyAsync(slots_dev + (size_t)f * slot_words, slots_host + (size_t)f * slot_words,
                                 (size_t)field_words[f] * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    }
    return CIA_OK;
}

int cia_rle_expand(cia_handle h, const uint32_t* slots_dev, int n_fields, size_t slot_words, int H, int W,
                   int32_t* labels_dev, void* stream) {
    if (!h) return CIA_E_ARG;
    if (!slots_dev || !labels_dev || H <= 0 || W <= 0) { h->err = "cia_rle_expand: bad argument"; return CIA_E_ARG; }
    if (n_fields <= 0) return CIA_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const dim3 grid((unsigned)((W + 1023) / 1024), (unsigned)H, (unsigned)n_fields);
    rle_expand_kernel<<<grid, 256, 0, s>>>(slots_dev, slot_words, H, W, labels_dev);
    CIA_LAUNCH_CHECK();
    return CIA_OK;
}

}  // extern "C"
