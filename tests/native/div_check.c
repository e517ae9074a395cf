/*
 * Exhaustive check of the division-free Clinger step used by ms_parse_next (test infrastructure).
 *
 *   a / 10^k  for every integer a in [1, 2^MS_DIV_BITS), MS_DIV_BITS = 26 by default (the tests), 32 checked once by hand and every k in [1, 22]
 *
 * computed as  q0 = a*y ; r0 = fma(-q0, d, a) ; q1 = fma(r0, y, q0)  [; r1 = fma(-q1, d, a) ; q2 = fma(r1, y, q1)]
 * with d = 10^k (exact) and y = RN(1/d), must equal the IEEE quotient bit for bit.
 *   gcc -O2 -mfma -fopenmp tests/native/div_check.c -o /tmp/div_check -lm && /tmp/div_check
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#ifndef MS_DIV_BITS
#define MS_DIV_BITS 26
#endif

int main(void) {
    static const double P10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                   1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    long long bad2 = 0, bad3 = 0;
    for (int k = 1; k <= 22; k++) {
        const double d = P10[k];
        const double y = 1.0 / d; /* correctly rounded reciprocal */
        long long b2 = 0, b3 = 0;
#pragma omp parallel for reduction(+ : b2, b3) schedule(static)
        for (long long ai = 1; ai < (1ll << MS_DIV_BITS); ai++) {
            const double a = (double)ai;
            const double want = a / d;
            const double q0 = a * y;
            const double r0 = __builtin_fma(-q0, d, a);
            const double q1 = __builtin_fma(r0, y, q0);
            const double r1 = __builtin_fma(-q1, d, a);
            const double q2 = __builtin_fma(r1, y, q1);
            uint64_t w, g1, g2;
            memcpy(&w, &want, 8);
            memcpy(&g1, &q1, 8);
            memcpy(&g2, &q2, 8);
            b2 += (g1 != w);
            b3 += (g2 != w);
        }
        printf("k=%2d  one-correction mismatches %lld  two-correction mismatches %lld\n", k, b2, b3);
        bad2 += b2;
        bad3 += b3;
    }
    printf("TOTAL one-correction %lld two-correction %lld\n", bad2, bad3);
    return bad3 != 0;
}
