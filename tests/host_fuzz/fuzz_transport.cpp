// ASan/UBSan fuzz harness for the multi-threaded host half of the label / image hand-over (csrc/transport.cu:
// cia_rle_encode_fields, cia_rle_encode_pack_fields -- the host work inside the timed region of the end-to-end
// bench): exact-size heap buffers, random field shapes, thread counts, slot and patch capacities (also too small
// ones: CIA_E_CAPACITY / "not packed" must be the answer, never a stray write), negative and out-of-range labels.
// Every accepted encoding is decoded and compared pixel by pixel; packed rectangles are recomputed from the labels.
// Built (nvcc, host functions only: no GPU is touched) and run by tests/test_host_fuzz.py;  fuzz_transport [iterations]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
extern "C" {
size_t cia_rle_slot_words(int H, int W);
int cia_rle_encode_fields(const int32_t*, int, int, int, uint32_t*, size_t, uint32_t*, int32_t*, int);
int cia_rle_encode_pack_fields(const int32_t*, const uint16_t*, int, int, int, uint32_t*, size_t, uint32_t*, int32_t*, int,
                               uint16_t*, size_t, uint32_t*, int);
}
int main(int argc, char** argv) {
    const int its = argc > 1 ? atoi(argv[1]) : 1500;
    std::mt19937_64 rng(99);
    long long ok = 0, cap_hits = 0, unpacked = 0;
    for (int it = 0; it < its; ++it) {
        int F = 1 + rng() % 6, H = 1 + rng() % 48, W = 1 + rng() % 120, T = 1 + rng() % 5;
        size_t px = (size_t)F * H * W;
        int32_t* lab = (int32_t*)malloc(px * 4); uint16_t* img = (uint16_t*)malloc(px * 2);
        int style = rng() % 3, nlab = 1 + rng() % 12;
        for (size_t i = 0; i < px; ++i) { img[i] = (uint16_t)rng();
            lab[i] = style == 0 ? (int32_t)(rng() % (nlab + 1)) : style == 1 ? ((rng() % 15) ? (i ? lab[i - 1] : 0) : (int32_t)(rng() % (nlab + 1))) : 0; }
        if (rng() % 7 == 0) lab[rng() % px] = -3;                 // negative labels are background
        if (rng() % 7 == 0) lab[rng() % px] = nlab + 50;          // above label_cap: skipped by the packer
        size_t sw = (rng() % 3) ? cia_rle_slot_words(H, W) : ((H + 2) & ~1) + 2 * (size_t)H * W;   // default slot or worst case
        if (rng() % 5 == 0) sw = rng() % (sw + 1);
        uint32_t* slots = (uint32_t*)malloc((F * sw ? F * sw : 1) * 4); uint32_t* fw = (uint32_t*)malloc(F * 4);
        size_t pcap = rng() % ((size_t)H * W * 2 + 1);
        uint16_t* patches = (uint16_t*)malloc((F * pcap ? F * pcap : 1) * 2); uint32_t* ppx = (uint32_t*)malloc(F * 4);
        int32_t mx = -1;
        int rc = (it & 1) ? cia_rle_encode_fields(lab, F, H, W, slots, sw, fw, &mx, T)
                          : cia_rle_encode_pack_fields(lab, img, F, H, W, slots, sw, fw, &mx, nlab, patches, pcap, ppx, T);
        if (rc == 0) {
            ++ok;
            for (int f = 0; f < F; ++f) {                         // decode every field and compare
                const uint32_t* slot = slots + (size_t)f * sw; const uint32_t* runs = slot + ((H + 2) & ~1);
                if (fw[f] > sw) { printf("words > slot\n"); return 1; }
                for (int y = 0; y < H; ++y) for (uint32_t k = slot[y]; k < slot[y + 1]; ++k) {
                    uint32_t x0 = runs[2 * k], x1 = k + 1 < slot[y + 1] ? runs[2 * k + 2] : (uint32_t)W;
                    for (uint32_t x = x0; x < x1; ++x) if (lab[((size_t)f * H + y) * W + x] != (int32_t)runs[2 * k + 1]) { printf("mismatch\n"); return 1; }
                }
                if (!(it & 1)) {
                    if (ppx[f] == 0xFFFFFFFFu) { ++unpacked; continue; }
                    if (ppx[f] > pcap) { printf("patch overflow\n"); return 1; }
                    // recompute boxes and compare the packed pixels
                    size_t n = 0;
                    for (int l = 1; l <= nlab; ++l) {
                        int r0 = H, r1 = 0, c0 = W, c1 = 0;
                        for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) if (lab[((size_t)f * H + y) * W + x] == l) { r0 = y < r0 ? y : r0; r1 = y + 1 > r1 ? y + 1 : r1; c0 = x < c0 ? x : c0; c1 = x + 1 > c1 ? x + 1 : c1; }
                        if (!r1) continue;
                        for (int y = r0; y < r1; ++y) for (int x = c0; x < c1; ++x) if (patches[(size_t)f * pcap + n++] != img[((size_t)f * H + y) * W + x]) { printf("patch pixel mismatch\n"); return 1; }
                    }
                    if (n != ppx[f]) { printf("patch count mismatch %zu %u\n", n, ppx[f]); return 1; }
                }
            }
        } else if (rc == -4) ++cap_hits; else { printf("rc %d\n", rc); return 1; }
        free(lab); free(img); free(slots); free(fw); free(patches); free(ppx);
    }
    printf("ok: %lld encoded, %lld capacity refusals, %lld fields left unpacked\n", ok, cap_hits, unpacked);
    return 0;
}
