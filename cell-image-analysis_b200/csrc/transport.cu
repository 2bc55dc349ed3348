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
#include "common.cuh"

#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>

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
