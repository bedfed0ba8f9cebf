/* glibc_math_core.h -- explicit-FMA restatement of the MAIN paths of glibc 2.39's double pow / exp / exp2 / log
 * (sysdeps/ieee754/dbl-64/e_pow.c, e_exp.c, e_exp2.c, e_log.c; Szabolcs Nagy's routines from ARM optimized-routines),
 * exactly as this image's libm.so.6 executes them on an FMA-capable x86-64: the ifunc'ed FMA builds of pow / exp / log
 * (GCC 13.3 contracts most a*b+c of the source into vfmadd; the pattern below was read off the disassembly of the same
 * source compiled with the same compiler, then verified bit-for-bit against the library on 50 M arguments each) and the
 * generic build of exp2 (no FMA variant exists).
 *
 * Why: the reference (Numba/LLVM) calls exactly these libm entry points, and every other fp64 operation
 * of the land kernels is IEEE-exact, so with these the CUDA path reproduces the reference BIT FOR BIT instead of to 1e-9.
 *
 * GLM_FMA(a,b,c) = fused a*b+c with one rounding; everything else is a plain IEEE operation (compile WITHOUT contraction:
 * nvcc -fmad=false / gcc -ffp-contract=off).  Each function returns 0 when the argument is outside the main path (zero,
 * negative, subnormal, inf, nan, |result exponent| >= 512 ...); the caller then uses another implementation, which is
 * exact for those special values.  Shared by the device build (glibc_math.cuh) and tests/test_glibc_math.py. */
#ifndef GLIBC_MATH_CORE_H
#define GLIBC_MATH_CORE_H
GLM_FN int glm_pow_main(double x, double y, double *out) {
    const uint64_t ix = glm_asuint64(x);
    const uint32_t topx = (uint32_t)(ix >> 52), topy = (uint32_t)(glm_asuint64(y) >> 52);
    if (topx - 0x001u >= 0x7ffu - 0x001u || (topy & 0x7ffu) - 0x3beu >= 0x43eu - 0x3beu) return 0;
    /* log_inline (e_pow.c) */
    const uint64_t tmp = ix - 0x3fe6955500000000ULL;
    const int i = (int)((tmp >> 45) & 127u);
    const int k = (int)((int64_t)tmp >> 52);
    const double z = glm_asdouble(ix - (tmp & (0xfffULL << 52)));
    const double kd = (double)k;
    const double invc = GLM_POWLOG_TAB(i, 0), logc = GLM_POWLOG_TAB(i, 2), logctail = GLM_POWLOG_TAB(i, 3);
    const double t1 = GLM_FMA(kd, GLM_POWLOG_LN2HI, logc);
    const double lo1 = GLM_FMA(kd, GLM_POWLOG_LN2LO, logctail);
    const double r = GLM_FMA(z, invc, -1.0);
    const double ar = r * GLM_POWLOG_A(0);
    const double q12 = GLM_FMA(r, GLM_POWLOG_A(2), GLM_POWLOG_A(1));
    const double q34 = GLM_FMA(r, GLM_POWLOG_A(4), GLM_POWLOG_A(3));
    const double t2 = r + t1;
    const double ar2 = r * ar;
    const double t1mt2 = t1 - t2;
    const double ar3 = r * ar2;
    const double lo3 = GLM_FMA(ar, r, -ar2);
    const double lo2 = t1mt2 + r;
    const double q56 = GLM_FMA(r, GLM_POWLOG_A(6), GLM_POWLOG_A(5));
    const double hi = t2 + ar2;
    const double t2mhi = t2 - hi;
    const double qa = GLM_FMA(q56, ar2, q34);
    const double lo4 = t2mhi + ar2;
    const double qb = GLM_FMA(ar2, qa, q12);
    const double lo = GLM_FMA(ar3, qb, ((lo1 + lo2) + lo3) + lo4);
    const double lhi = hi + lo;
    const double ehi = y * lhi;
    const uint32_t abstop = (uint32_t)(glm_asuint64(ehi) >> 52) & 0x7ffu;
    if (abstop - 0x3c9u >= 0x408u - 0x3c9u) return 0; /* |ehi| outside [2^-54, 2^9) */
    /* exp_inline */
    const double kd1 = GLM_FMA(ehi, GLM_EXP_INVLN2N, GLM_EXP_SHIFT);
    const double llo = (hi - lhi) + lo;
    const uint64_t ki = glm_asuint64(kd1);
    const double kd2 = kd1 - GLM_EXP_SHIFT;
    const double ra = GLM_FMA(kd2, GLM_EXP_NEGLN2HIN, ehi);
    const double e1 = GLM_FMA(y, lhi, -ehi);
    const double rb = GLM_FMA(kd2, GLM_EXP_NEGLN2LON, ra);
    const double elo = GLM_FMA(y, llo, e1);
    const double rr = rb + elo;
    const int idx = 2 * (int)(ki & 127u);
    const double p23 = GLM_FMA(rr, GLM_EXP_C(1), GLM_EXP_C(0));
    const double tr = rr + glm_asdouble(GLM_EXP_TAB(idx));
    const double r2 = rr * rr;
    const double p45 = GLM_FMA(rr, GLM_EXP_C(3), GLM_EXP_C(2));
    const double s1 = GLM_FMA(p23, r2, tr);
    const double r4 = r2 * r2;
    const double tmpv = GLM_FMA(p45, r4, s1);
    const double scale = glm_asdouble(GLM_EXP_TAB(idx + 1) + (ki << 45));
    *out = GLM_FMA(scale, tmpv, scale);
    return 1;
}

GLM_FN int glm_exp_main(double x, double *out) {
    const uint32_t abstop = (uint32_t)(glm_asuint64(x) >> 52) & 0x7ffu;
    if (abstop - 0x3c9u >= 0x408u - 0x3c9u) return 0;
    const double kd1 = GLM_FMA(x, GLM_EXP_INVLN2N, GLM_EXP_SHIFT);
    const uint64_t ki = glm_asuint64(kd1);
    const double kd2 = kd1 - GLM_EXP_SHIFT;
    const double ra = GLM_FMA(kd2, GLM_EXP_NEGLN2HIN, x);
    const double rr = GLM_FMA(kd2, GLM_EXP_NEGLN2LON, ra);
    const int idx = 2 * (int)(ki & 127u);
    const double p23 = GLM_FMA(rr, GLM_EXP_C(1), GLM_EXP_C(0));
    const double tr = rr + glm_asdouble(GLM_EXP_TAB(idx));
    const double r2 = rr * rr;
    const double p45 = GLM_FMA(rr, GLM_EXP_C(3), GLM_EXP_C(2));
    const double s1 = GLM_FMA(p23, r2, tr);
    const double r4 = r2 * r2;
    const double tmpv = GLM_FMA(p45, r4, s1);
    const double scale = glm_asdouble(GLM_EXP_TAB(idx + 1) + (ki << 45));
    *out = GLM_FMA(scale, tmpv, scale);
    return 1;
}

/* exp2: the generic (non-FMA) build, e_exp2.c evaluated left to right */
GLM_FN int glm_exp2_main(double x, double *out) {
    const uint32_t abstop = (uint32_t)(glm_asuint64(x) >> 52) & 0x7ffu;
    if (abstop - 0x3c9u >= 0x408u - 0x3c9u) return 0;
    const double kd1 = x + GLM_EXP2_SHIFT;
    const uint64_t ki = glm_asuint64(kd1);
    const double kd2 = kd1 - GLM_EXP2_SHIFT;
    const double r = x - kd2;
    const int idx = 2 * (int)(ki & 127u);
    const double tail = glm_asdouble(GLM_EXP_TAB(idx));
    const double r2 = r * r;
    const double tmpv = ((tail + r * GLM_EXP2_C(0)) + r2 * (GLM_EXP2_C(1) + r * GLM_EXP2_C(2))) +
                        (r2 * r2) * (GLM_EXP2_C(3) + r * GLM_EXP2_C(4));
    const double scale = glm_asdouble(GLM_EXP_TAB(idx + 1) + (ki << 45));
    *out = scale + scale * tmpv;
    return 1;
}

GLM_FN int glm_log_main(double x, double *out) {
    const uint64_t ix = glm_asuint64(x);
    const uint64_t LO = 0x3fee000000000000ULL, HI = 0x3ff1090000000000ULL; /* 1 - 2^-4, 1 + 0x1.09p-4 */
    if (ix - LO < HI - LO) { /* close to 1 */
        if (ix == 0x3ff0000000000000ULL) {
            *out = 0.0;
            return 1;
        }
        const double r = x - 1.0;
        const double a12 = GLM_FMA(r, GLM_LOG_B(2), GLM_LOG_B(1));
        const double a78 = GLM_FMA(r, GLM_LOG_B(8), GLM_LOG_B(7));
        const double r2 = r * r;
        const double a45 = GLM_FMA(r, GLM_LOG_B(5), GLM_LOG_B(4));
        const double q1 = GLM_FMA(r2, GLM_LOG_B(3), a12);
        const double q2 = GLM_FMA(r2, GLM_LOG_B(6), a45);
        const double r3 = r * r2;
        const double a79 = GLM_FMA(r2, GLM_LOG_B(9), a78);
        const double q3 = GLM_FMA(r3, GLM_LOG_B(10), a79);
        const double q23 = GLM_FMA(q3, r3, q2);
        const double Q = GLM_FMA(q23, r3, q1);
        const double rw = GLM_FMA(r, 0x1p27, r);
        const double rhi = GLM_FMA(-0x1p27, r, rw);
        const double rhi2 = rhi * rhi;
        const double rlo = r - rhi;
        const double hi = GLM_FMA(rhi2, GLM_LOG_B(0), r);
        const double rmhi = r - hi;
        const double rprhi = r + rhi;
        const double lo = GLM_FMA(rhi2, GLM_LOG_B(0), rmhi);
        const double b0rlo = GLM_LOG_B(0) * rlo;
        const double lo2 = GLM_FMA(b0rlo, rprhi, lo);
        const double y = GLM_FMA(Q, r3, lo2);
        *out = hi + y;
        return 1;
    }
    const uint32_t top = (uint32_t)(ix >> 48);
    if (top - 0x0010u >= 0x7ff0u - 0x0010u) return 0;
    const uint64_t tmp = ix - 0x3fe6000000000000ULL;
    const int i = (int)((tmp >> 45) & 127u);
    const int k = (int)((int64_t)tmp >> 52);
    const double z = glm_asdouble(ix - (tmp & (0xfffULL << 52)));
    const double kd = (double)k;
    const double w = GLM_FMA(kd, GLM_LOG_LN2HI, GLM_LOG_TAB(i, 1));
    const double r = GLM_FMA(z, GLM_LOG_TAB(i, 0), -1.0);
    const double a12 = GLM_FMA(r, GLM_LOG_A(2), GLM_LOG_A(1));
    const double hi = r + w;
    const double r2 = r * r;
    const double lo = GLM_FMA(kd, GLM_LOG_LN2LO, (w - hi) + r);
    const double r3 = r * r2;
    const double a34 = GLM_FMA(r, GLM_LOG_A(4), GLM_LOG_A(3));
    const double l0 = GLM_FMA(r2, GLM_LOG_A(0), lo);
    const double pq = GLM_FMA(a34, r2, a12);
    const double yy = GLM_FMA(pq, r3, l0);
    *out = yy + hi;
    return 1;
}
#endif
