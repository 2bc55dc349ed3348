/* The projection kernel (score.cu, store_x) computes RobustScaler's v / scale as
 *   q = v * r;  rem = fma(-q, scale, v);  q' = fma(rem, r, q)      with r = 1.0 / scale (host, IEEE)
 * instead of an IEEE division.  This program checks q' == v / scale bit for bit (and therefore
 * the float32 rounding that follows) over random float32 numerators and double scales of many
 * magnitudes, plus scales whose significands are all ones / powers of two.  Compile with
 * -ffp-contract=off so that only the explicit fma() calls fuse. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static uint64_t s_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd(void) {
    s_state ^= s_state << 13; s_state ^= s_state >> 7; s_state ^= s_state << 17;
    return s_state;
}
static float rnd_float(void) {           /* finite float32 with exponent in [-30, 30] */
    uint32_t m = (uint32_t)rnd() & 0x7FFFFFu, e = 97u + (uint32_t)(rnd() % 61u), sg = (uint32_t)(rnd() & 1u);
    uint32_t u = (sg << 31) | (e << 23) | m;
    float f; memcpy(&f, &u, 4); return f;
}
static double rnd_scale(int mode) {      /* positive double, exponent in [-20, 20] */
    uint64_t m = rnd() & 0xFFFFFFFFFFFFFull, e = 1003ull + rnd() % 41ull;
    if (mode == 1) m = 0xFFFFFFFFFFFFFull;            /* significand all ones */
    if (mode == 2) m = 0;                              /* power of two */
    if (mode == 3) m &= 0xFFFFFE0000000ull;            /* a float32 value (scale_ fit on float32 data) */
    uint64_t u = (e << 52) | m;
    double d; memcpy(&d, &u, 8); return d;
}

int main(void) {
    long n = 0, bad = 0;
    for (int mode = 0; mode < 4; ++mode) {
        const long scales = mode == 0 ? 20000 : 4000;
        for (long i = 0; i < scales; ++i) {
            const double s = rnd_scale(mode), r = 1.0 / s;
            for (int k = 0; k < 1000; ++k) {
                const double a = (double)rnd_float();
                const double q = a * r;
                const double rem = fma(-q, s, a);
                const double q2 = fma(rem, r, q);
                const double ref = a / s;
                ++n;
                if (memcmp(&q2, &ref, 8) != 0 || (float)q2 != (float)ref) ++bad;
            }
        }
    }
    printf("%ld quotients, %ld mismatches\n", n, bad);
    return bad != 0;
}
