// const_div.h -- a / C for a compile-time constant C = m * 2^k whose odd part m is a SMALL integer (24 = 3*2^3, 100 = 25*2^2,
// 144 = 9*2^4, ...): the correctly rounded quotient in three fp64 operations instead of the ~20 instructions and ~8 dependent
// operations of the IEEE division sequence.  Same bits as a / C for every argument:
//
//   y  = RN(1/C)            (folded by the compiler: an IEEE division of two constants)
//   q  = RN(a*y)            |q - a/C| <= 2 ulp
//   r  = fma(-C, q, a)      EXACT: a - C*q is a multiple of 2^k*ulp(q)/4 of magnitude <= ~8m such units
//   q' = fma(r, y, q)     = RN(a/C + d),  |d| = |r| * |y - 1/C| <= 2^-51.9 ulp
//
// and a/C can never sit closer than 1/(2m) ulp to a rounding boundary: in units of its own ulp it is N/m for an integer N, a
// boundary is j + 1/2, and |N/m - j - 1/2| = |2N - m(2j+1)| / (2m) >= 1/(2m) because m(2j+1) is odd.  So RN(a/C + d) = RN(a/C)
// whenever 2^-51.9 < 1/(2m), i.e. for every m below 2^50 (m <= 1815 here).  The argument needs no overflow, underflow or special
// value anywhere: arguments outside 2^-900 <= |a| < 2^900 (zeros, subnormals, inf, nan among them) take the ordinary division.
// tests/const_div_check.c pins the sequence against `/` on the host (explicit fma, no contraction), tests/test_glibc_math.py.
#pragma once
#ifndef HSP_CDIV_FMA
#error "define HSP_CDIV_FMA(a, b, c), HSP_CDIV_HI(x) (high 32 bits of a double) and HSP_CDIV_SLOW(a, c) before including const_div.h"
#endif
#define HSP_CDIV_FN(name, C)                                                      \
    HSP_CDIV_DECL double name(double a) {                                         \
        const double c = (C), y = 1.0 / (C);                                      \
        const unsigned e = ((unsigned)HSP_CDIV_HI(a) >> 20) & 0x7ffu;             \
        if (e - 123u < 1800u) { /* 2^-900 <= |a| < 2^900 */                       \
            const double q = a * y;                                               \
            const double r = HSP_CDIV_FMA(-c, q, a);                              \
            return HSP_CDIV_FMA(r, y, q);                                         \
        }                                                                         \
        return HSP_CDIV_SLOW(a, c);                                               \
    }
HSP_CDIV_FN(hsp_div24, 24.0)
HSP_CDIV_FN(hsp_div100, 100.0)
HSP_CDIV_FN(hsp_div144, 144.0)
HSP_CDIV_FN(hsp_div10, 10.0)
HSP_CDIV_FN(hsp_div1440, 1440.0)
HSP_CDIV_FN(hsp_div3630, 3630.0)
