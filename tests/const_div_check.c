/* tests/const_div_check.c -- TEST: hspsquared_b200/csrc/const_div.h (division by a small-odd-part constant in three operations)
 * against the host's IEEE division, explicit fma() and no contraction.  usage: const_div_check N -> mismatch counts, 0 = pass. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static inline uint32_t hi32(double f) { uint64_t u; memcpy(&u, &f, 8); return (uint32_t)(u >> 32); }
static inline uint64_t bits(double f) { uint64_t u; memcpy(&u, &f, 8); return u; }
static inline double from_bits(uint64_t u) { double f; memcpy(&f, &u, 8); return f; }
#define HSP_CDIV_DECL static inline
#define HSP_CDIV_FMA(a, b, c) fma((a), (b), (c))
#define HSP_CDIV_HI(x) hi32(x)
#define HSP_CDIV_SLOW(a, c) ((a) / (c))
#include "const_div.h"

static uint64_t s = 88172645463325252ULL;
static uint64_t rnd64(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }

static long check(const char *name, double (*f)(double), double c, long n) {
    long bad = 0;
    volatile double cv = c; /* a run-time divisor: the compiler must emit a real division */
    for (long i = 0; i < n; i++) {
        uint64_t u = rnd64();
        double a;
        switch (i % 5) {
        case 0: a = from_bits(u); break; /* any bit pattern: specials, subnormals, huge, tiny */
        case 1: a = from_bits((u & 0x800fffffffffffffULL) | ((uint64_t)(1023 - 40 + (int)(rnd64() % 80)) << 52)); break;
        case 2: { /* exact multiples of c and their neighbours: quotients that are exactly representable / one ulp off */
            double qx = from_bits((u & 0x000fffffffffffffULL) | ((uint64_t)(1023 - 20 + (int)(rnd64() % 40)) << 52));
            a = qx * c;
            if (rnd64() & 1) a = nextafter(a, (rnd64() & 1) ? INFINITY : -INFINITY);
            break;
        }
        case 3: a = (double)(int64_t)(rnd64() % 2000000) / 1000.0; break; /* short decimals, as model values often are */
        default: a = from_bits((u & 0x800fffffffffffffULL) | ((uint64_t)(1 + (int)(rnd64() % 2046)) << 52)); break; /* all normal exponents */
        }
        const double g = f(a), w = a / cv;
        if (bits(g) != bits(w) && !(g != g && w != w)) {
            if (bad < 5) printf("%s: a=%a got %a want %a\n", name, a, g, w);
            bad++;
        }
    }
    printf("%s %ld/%ld\n", name, bad, n);
    return bad;
}

int main(int argc, char **argv) {
    long n = argc > 1 ? atol(argv[1]) : 1000000, bad = 0;
    bad += check("div24", hsp_div24, 24.0, n);
    bad += check("div100", hsp_div100, 100.0, n);
    bad += check("div144", hsp_div144, 144.0, n);
    bad += check("div10", hsp_div10, 10.0, n);
    bad += check("div1440", hsp_div1440, 1440.0, n);
    bad += check("div3630", hsp_div3630, 3630.0, n);
    return bad != 0;
}
