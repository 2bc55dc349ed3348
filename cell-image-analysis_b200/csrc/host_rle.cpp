// host_rle.cpp -- host side of the run-length label transport (see transport.cu): one pass
// over an int32 label field at memory speed.  Plain C++ (compiled by g++, no CUDA); the inner
// "how long does this run last" scan has an AVX2 form chosen at run time.
#include <cstddef>
#include <cstdint>
#include <cstring>

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define CIA_HAVE_AVX2_PATH 1
#else
#define CIA_HAVE_AVX2_PATH 0
#endif

namespace {

inline size_t runs_base(int H) { return (size_t)((H + 2) & ~1); }

// first index >= x with p[index] != cur (or W)
inline int run_end_scalar(const int32_t* p, int x, int W, int32_t cur) {
    const uint64_t pat = (uint64_t)(uint32_t)cur * 0x0000000100000001ULL;
    while (x + 8 <= W) {                      // 8 labels (32 bytes) per step while the run lasts
        uint64_t v[4];
        std::memcpy(v, p + x, 32);
        if (((v[0] ^ pat) | (v[1] ^ pat) | (v[2] ^ pat) | (v[3] ^ pat)) != 0) break;
        x += 8;
    }
    while (x < W && p[x] == cur) ++x;
    return x;
}

#if CIA_HAVE_AVX2_PATH
__attribute__((target("avx2"))) inline int run_end_avx2(const int32_t* p, int x, int W, int32_t cur) {
    const __m256i pat = _mm256_set1_epi32(cur);
    while (x + 32 <= W) {                     // 32 labels per step
        const __m256i c0 = _mm256_cmpeq_epi32(_mm256_loadu_si256((const __m256i*)(p + x)), pat);
        const __m256i c1 = _mm256_cmpeq_epi32(_mm256_loadu_si256((const __m256i*)(p + x + 8)), pat);
        const __m256i c2 = _mm256_cmpeq_epi32(_mm256_loadu_si256((const __m256i*)(p + x + 16)), pat);
        const __m256i c3 = _mm256_cmpeq_epi32(_mm256_loadu_si256((const __m256i*)(p + x + 24)), pat);
        if (_mm256_movemask_epi8(_mm256_and_si256(_mm256_and_si256(c0, c1), _mm256_and_si256(c2, c3))) != -1) break;
        x += 32;
    }
    while (x + 8 <= W) {
        const int m = _mm256_movemask_ps(_mm256_castsi256_ps(
            _mm256_cmpeq_epi32(_mm256_loadu_si256((const __m256i*)(p + x)), pat)));
        if (m != 0xFF) return x + __builtin_ctz(~m & 0xFF);
        x += 8;
    }
    while (x < W && p[x] == cur) ++x;
    return x;
}
#endif

template <class RunEnd>
inline size_t encode_field_impl(const int32_t* lab, int H, int W, uint32_t* slot, size_t slot_words,
                                int32_t* max_label, RunEnd run_end) {
    const size_t r0 = runs_base(H);
    if (slot_words < r0 + 2) return 0;
    uint32_t* runs = slot + r0;
    const size_t cap_runs = (slot_words - r0) / 2;
    size_t n = 0;
    int32_t mx = 0;
    if ((H + 1) & 1) slot[H + 1] = 0;   // padding word
    for (int y = 0; y < H; ++y) {
        const int32_t* p = lab + (size_t)y * W;
        slot[y] = (uint32_t)n;
        const bool tight = n + (size_t)W > cap_runs;      // near the end of the slot: check every emit
        int x = 0;
        while (x < W) {
            const int32_t cur = p[x];
            if (tight && n >= cap_runs) return 0;
            runs[2 * n] = (uint32_t)x; runs[2 * n + 1] = (uint32_t)cur; ++n;
            if (cur > mx) mx = cur;
            x = run_end(p, x + 1, W, cur);
        }
    }
    slot[H] = (uint32_t)n;
    if (max_label) *max_label = mx;
    return r0 + 2 * n;
}

#if CIA_HAVE_AVX2_PATH
__attribute__((target("avx2"))) size_t encode_field_avx2(const int32_t* lab, int H, int W, uint32_t* slot,
                                                         size_t slot_words, int32_t* max_label) {
    return encode_field_impl(lab, H, W, slot, slot_words, max_label,
                             [](const int32_t* p, int x, int w, int32_t cur) __attribute__((target("avx2"))) {
                                 return run_end_avx2(p, x, w, cur);
                             });
}
#endif

}  // namespace

// Encodes one field; returns the number of words used, or 0 if the slot is too small.
size_t cia_host_encode_field(const int32_t* lab, int H, int W, uint32_t* slot, size_t slot_words,
                             int32_t* max_label) {
#if CIA_HAVE_AVX2_PATH
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) return encode_field_avx2(lab, H, W, slot, slot_words, max_label);
#endif
    return encode_field_impl(lab, H, W, slot, slot_words, max_label, run_end_scalar);
}
