// host_rle.cpp -- host side of the run-length label transport (see transport.cu): one pass
// over an int32 label field at memory speed.  Plain C++ (compiled by g++, no CUDA); the AVX2
// form (chosen at run time) compares 32 labels with their left neighbours per step.
#include <cstddef>
#include <cstdint>
#include <cstdlib>
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
// AVX2 form: instead of asking "how long does this run last" run by run (a restart of the wide
// compare loop at every boundary -- ~19 runs per 2048-pixel row cost 28 % of the scan speed), every
// 32 labels are compared with their left neighbours in one streaming step and the "differs from
// the previous pixel" bits are walked with ctz: a run starts exactly at every set bit.  Same
// output, word for word, as the scalar path.
__attribute__((target("avx2"))) size_t encode_field_avx2(const int32_t* lab, int H, int W, uint32_t* slot,
                                                         size_t slot_words, int32_t* max_label) {
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
#define CIA_EMIT(xx)                                                        \
        do {                                                                \
            if (tight && n >= cap_runs) return 0;                           \
            const int32_t l_ = p[(xx)];                                     \
            runs[2 * n] = (uint32_t)(xx); runs[2 * n + 1] = (uint32_t)l_; ++n; \
            if (l_ > mx) mx = l_;                                           \
        } while (0)
        CIA_EMIT(0);
        int x = 1;
        for (; x + 32 <= W; x += 32) {
            const __m256i e0 = _mm256_cmpeq_epi32(_mm256_loadu_si256((const __m256i*)(p + x)),
                                                  _mm256_loadu_si256((const __m256i*)(p + x - 1)));
            const __m256i e1 = _mm256_cmpeq_epi32(_mm256_loadu_si256((const __m256i*)(p + x + 8)),
                                                  _mm256_loadu_si256((const __m256i*)(p + x + 7)));
            const __m256i e2 = _mm256_cmpeq_epi32(_mm256_loadu_si256((const __m256i*)(p + x + 16)),
                                                  _mm256_loadu_si256((const __m256i*)(p + x + 15)));
            const __m256i e3 = _mm256_cmpeq_epi32(_mm256_loadu_si256((const __m256i*)(p + x + 24)),
                                                  _mm256_loadu_si256((const __m256i*)(p + x + 23)));
            // pack the four 8 x 32-bit compare results into one 32-bit "equal" mask (bit k = label x + k)
            const __m256i p01 = _mm256_packs_epi32(e0, e1);            // lanes: e0[0..3] e1[0..3] | e0[4..7] e1[4..7]
            const __m256i p23 = _mm256_packs_epi32(e2, e3);
            __m256i pk = _mm256_packs_epi16(p01, p23);                 // bytes, lane-interleaved
            pk = _mm256_permutevar8x32_epi32(pk, _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7));
            uint32_t ne = ~(uint32_t)_mm256_movemask_epi8(pk);
            while (ne) {
                const int k = __builtin_ctz(ne);
                ne &= ne - 1;
                CIA_EMIT(x + k);
            }
        }
        for (; x < W; ++x)
            if (p[x] != p[x - 1]) CIA_EMIT(x);
#undef CIA_EMIT
    }
    slot[H] = (uint32_t)n;
    if (max_label) *max_label = mx;
    return r0 + 2 * n;
}
#endif

}  // namespace

// Encodes one field; returns the number of words used, or 0 if the slot is too small.
size_t cia_host_encode_field(const int32_t* lab, int H, int W, uint32_t* slot, size_t slot_words,
                             int32_t* max_label) {
#if CIA_HAVE_AVX2_PATH
    // CIA_HOST_RLE_SCALAR=1 forces the portable path (tests compare the two word for word)
    static const bool avx2 = __builtin_cpu_supports("avx2") && getenv("CIA_HOST_RLE_SCALAR") == nullptr;
    if (avx2) return encode_field_avx2(lab, H, W, slot, slot_words, max_label);
#endif
    return encode_field_impl(lab, H, W, slot, slot_words, max_label, run_end_scalar);
}
