// ASan/UBSan fuzz harness for the host-only C++ of libcia (csrc/host_tiff.cpp: the strip / tile decoders behind
// tiff.imread, improved_detection.py:51; csrc/host_rle.cpp: the run-length encoder of the label hand-over): valid
// streams from an independent libtiff-flavoured LZW encoder must round-trip exactly, clipped capacities must clip,
// corrupted / truncated / random streams must neither crash nor write out of bounds (every buffer is an exact-size
// heap block so that AddressSanitizer traps the first stray byte).  Built and run by tests/test_host_fuzz.py;
//     fuzz_host [lzw iterations] [rle iterations]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
extern "C" long long cia_tiff_lzw_decode(const uint8_t*, size_t, uint8_t*, size_t);
extern "C" long long cia_tiff_packbits_decode(const uint8_t*, size_t, uint8_t*, size_t);
size_t cia_host_encode_field(const int32_t* lab, int H, int W, uint32_t* slot, size_t slot_words, int32_t* max_label);

// reference LZW encoder (TIFF flavour, MSB first, early change) to make valid streams
static std::vector<uint8_t> lzw_encode(const std::vector<uint8_t>& in) {
    std::vector<uint8_t> out; uint64_t acc = 0; int have = 0;
    auto put = [&](int code, int bits) { acc = (acc << bits) | (uint32_t)code; have += bits; while (have >= 8) { out.push_back((uint8_t)(acc >> (have - 8))); have -= 8; } };
    std::vector<int> tab(4096 * 256, -1);
    int next = 258, bits = 9; put(256, bits);
    int cur = -1;
    for (uint8_t c : in) {
        if (cur < 0) { cur = c; continue; }
        int& e = tab[(size_t)cur * 256 + c];
        if (e >= 0) { cur = e; continue; }
        put(cur, bits);
        e = next++;
        if (next >= (1 << bits) && bits < 12) ++bits;         // libtiff: free_ent > MAXCODE(nbits)
        if (next >= 4094) { put(256, bits); std::fill(tab.begin(), tab.end(), -1); next = 258; bits = 9; }
        cur = c;
    }
    if (cur >= 0) put(cur, bits);
    put(257, bits);
    if (have) out.push_back((uint8_t)(acc << (8 - have)));
    return out;
}

int main(int argc, char** argv) {
    const int lzw_its = argc > 1 ? atoi(argv[1]) : 20000, rle_its = argc > 2 ? atoi(argv[2]) : 3000;
    std::mt19937_64 rng(1234);
    long long checks = 0, roundtrips = 0;
    for (int it = 0; it < lzw_its; ++it) {
        size_t n = (it % 50 == 0) ? rng() % 40000 : rng() % 600;
        std::vector<uint8_t> plain(n);
        int mode = rng() % 3;
        for (auto& b : plain) b = mode == 0 ? (uint8_t)rng() : mode == 1 ? (uint8_t)(rng() % 3) : (uint8_t)((rng() % 50) ? 7 : rng());
        std::vector<uint8_t> enc = lzw_encode(plain);
        {   // valid stream, exact capacity
            uint8_t* src = (uint8_t*)malloc(enc.size() ? enc.size() : 1); if (!enc.empty()) memcpy(src, enc.data(), enc.size());
            uint8_t* dst = (uint8_t*)malloc(n ? n : 1);
            long long r = cia_tiff_lzw_decode(src, enc.size(), dst, n);
            if (r == (long long)n && (n == 0 || memcmp(dst, plain.data(), n) == 0)) ++roundtrips;
            else if (n) { printf("LZW round trip FAILED it=%d n=%zu r=%lld\n", it, n, r); return 1; }
            // short capacity: must clip, not overflow
            size_t cap = n ? rng() % n : 0;
            uint8_t* d2 = (uint8_t*)malloc(cap ? cap : 1);
            r = cia_tiff_lzw_decode(src, enc.size(), d2, cap);
            if (r > (long long)cap) { printf("LZW overflowed cap\n"); return 1; }
            free(d2);
            // corrupted stream
            for (int k = 0; k < 4 && !enc.empty(); ++k) src[rng() % enc.size()] ^= (uint8_t)(1u << (rng() % 8));
            size_t cut = enc.size() ? rng() % (enc.size() + 1) : 0;
            r = cia_tiff_lzw_decode(src, cut, dst, n);
            if (r > (long long)n) { printf("LZW overflowed on corrupt input\n"); return 1; }
            free(src); free(dst); ++checks;
        }
        {   // random bytes as a stream
            size_t m = rng() % 300, cap = rng() % 2000;
            uint8_t* src = (uint8_t*)malloc(m ? m : 1); for (size_t i = 0; i < m; ++i) src[i] = (uint8_t)rng();
            uint8_t* dst = (uint8_t*)malloc(cap ? cap : 1);
            long long r = cia_tiff_lzw_decode(src, m, dst, cap);
            if (r > (long long)cap) return 1;
            r = cia_tiff_packbits_decode(src, m, dst, cap);
            if (r > (long long)cap) return 1;
            free(src); free(dst); ++checks;
        }
    }
    // run-length encoder: exact-size label fields and slots of every capacity around the need
    for (int it = 0; it < rle_its; ++it) {
        int H = 1 + rng() % 40, W = 1 + rng() % 150;
        int32_t* lab = (int32_t*)malloc(sizeof(int32_t) * H * W);
        int style = rng() % 3;
        for (int i = 0; i < H * W; ++i) lab[i] = style == 0 ? (int32_t)(rng() % 5) : style == 1 ? ((rng() % 20) ? (i ? lab[i - 1] : 0) : (int32_t)(rng() % 1000)) : 0;
        size_t need = ((H + 2) & ~1) + 2 * (size_t)H * W;
        size_t words = rng() % (need + 8);
        uint32_t* slot = (uint32_t*)malloc(sizeof(uint32_t) * (words ? words : 1));
        int32_t mx = -1;
        size_t used = cia_host_encode_field(lab, H, W, slot, words, &mx);
        if (used > words) { printf("RLE used more than the slot\n"); return 1; }
        if (used) {   // decode and compare
            size_t r0 = (H + 2) & ~1; const uint32_t* runs = slot + r0;
            for (int y = 0; y < H; ++y) {
                uint32_t a = slot[y], b = y + 1 < H ? slot[y + 1] : slot[H];
                for (uint32_t k = a; k < b; ++k) {
                    uint32_t x0 = runs[2 * k], x1 = k + 1 < b ? runs[2 * k + 2] : (uint32_t)W;
                    for (uint32_t x = x0; x < x1; ++x) if (lab[y * W + x] != (int32_t)runs[2 * k + 1]) { printf("RLE mismatch\n"); return 1; }
                }
            }
            ++roundtrips;
        }
        free(lab); free(slot); ++checks;
    }
    printf("ok: %lld checks, %lld exact round trips\n", checks, roundtrips);
    return 0;
}
