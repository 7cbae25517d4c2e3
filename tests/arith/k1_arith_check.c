/* tests/arith/k1_arith_check.c -- checks the exactly-rounded constant-divisor sequences of
 * uncalled_b200/csrc/unc_k1.cuh (k1_fdiv_w / k1_ddiv_w) against IEEE division:
 *   float : EXHAUSTIVELY over all 2^32 bit patterns, divisors 3 and 6
 *   double: N random operands (all exponents, plus patterns near powers of two), divisors 3 and 6
 * Build: gcc -O2 -mfma -ffp-contract=off -fopenmp k1_arith_check.c -lm ; prints "OK" or mismatches. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* as k1_fdiv_w: below 2^-100 the device falls back to the IEEE division (ties exist among subnormal quotients) */
static inline float fdiv_w(float a, float w, float r) { if (fabsf(a) < 0x1p-100f) return a / w; float q = a * r; float e = fmaf(-w, q, a); return fmaf(e, r, q); }
static inline double ddiv_w(double a, double w, double r) { double q = a * r; double e = fma(-w, q, a); return fma(e, r, q); }

static inline uint64_t rng(uint64_t *s) { *s ^= *s << 13; *s ^= *s >> 7; *s ^= *s << 17; return *s; }

int main(int argc, char **argv) {
    long nd = argc > 1 ? atol(argv[1]) : 200000000L;
    long bad = 0;
    const float fw[2] = {3.0f, 6.0f}, fr[2] = {0.333333343f, 0.166666672f};
    for (int k = 0; k < 2; k++) {
#pragma omp parallel for reduction(+ : bad) schedule(static)
        for (long long b = 0; b < (1LL << 32); b++) {
            uint32_t u = (uint32_t) b; float a; memcpy(&a, &u, 4);
            if (a != a || isinf(a)) continue;
            float want = a / fw[k], got = fdiv_w(a, fw[k], fr[k]);
            uint32_t x, y; memcpy(&x, &want, 4); memcpy(&y, &got, 4);
            if (x != y && !(want == 0.0f && got == 0.0f)) { if (bad < 10) fprintf(stderr, "float /%g: a=%a want=%a got=%a\n", fw[k], a, want, got); bad++; }
        }
    }
    const double dw[2] = {3.0, 6.0}, dr[2] = {0.33333333333333331, 0.16666666666666666};
    for (int k = 0; k < 2; k++) {
#pragma omp parallel for reduction(+ : bad) schedule(static)
        for (long i = 0; i < nd; i++) {
            uint64_t s = 0x9E3779B97F4A7C15ull * (uint64_t) (i + 1) + 12345u + (uint64_t) k;
            uint64_t v = rng(&s);
            if ((i & 7) == 0) v = (v & 0xFFF0000000000000ull) | (rng(&s) & 0xFF);              /* just above a power of two */
            if ((i & 7) == 1) v = (v & 0xFFF0000000000000ull) | (0xFFFFFFFFFFFFFull - (rng(&s) & 0xFF)); /* just below */
            double a; memcpy(&a, &v, 8);
            int ex = (int) ((v >> 52) & 0x7FF);
            if (ex < 60 || ex > 1990) continue;               /* the device uses the sequence far from over/underflow */
            double want = a / dw[k], got = ddiv_w(a, dw[k], dr[k]);
            uint64_t x, y; memcpy(&x, &want, 8); memcpy(&y, &got, 8);
            if (x != y) { if (bad < 10) fprintf(stderr, "double /%g: a=%a want=%a got=%a\n", dw[k], a, want, got); bad++; }
        }
    }
    if (bad) { printf("MISMATCHES %ld\n", bad); return 1; }
    printf("OK\n");
    return 0;
}
