#include <math.h>
#include <stdio.h>
int main(void) {
    long bad = 0, total = 0;
    for (int den = 1; den <= 16383; ++den) {
        const double d = (double)den, y = 1.0 / d;
        for (int a = 0; a <= den; ++a) {
            const double x = (double)a;
            const double q0 = x * y;
            const double rem = fma(-q0, d, x);
            const double q1 = fma(rem, y, q0);
            if (q1 != x / d) { if (bad < 5) printf("mismatch %d/%d\n", a, den); ++bad; }
            ++total;
        }
    }
    printf("checked %ld quotients, %ld mismatches\n", total, bad);
    return bad != 0;
}
