// glibc_math.cuh -- device pow / exp / exp2 / log that return the very bits glibc 2.39's libm returns (see
// glibc_math_core.h).  Arguments outside the main path (zero, negative, subnormal, inf, nan, huge exponents) go to
// CUDA's libm, whose results for those special values are exact.  The test double (host build) calls glibc itself.
#pragma once
#include "cuda_rt.h"
#include <stdint.h>

#ifdef HSP2G_HOST_EMU
#include <cmath>
static inline double hsp_pow(double x, double y) { return pow(x, y); }
static inline double hsp_exp(double x) { return exp(x); }
static inline double hsp_exp2(double x) { return exp2(x); }
static inline double hsp_log(double x) { return log(x); }
#else
#define GLM_TABLE static __device__ const
#include "glibc_tables.h"
#define GLM_FN static __device__ __forceinline__
#define GLM_FMA(a, b, c) __fma_rn((a), (b), (c))
#define glm_asuint64(x) ((uint64_t)__double_as_longlong(x))
#define glm_asdouble(u) __longlong_as_double((long long)(u))
#define GLM_POWLOG_A(j) GLM_POWLOG_A##j
#define GLM_EXP_C(j) GLM_EXP_C##j
#define GLM_EXP2_C(j) GLM_EXP2_C##j
#define GLM_LOG_A(j) GLM_LOG_A##j
#define GLM_LOG_B(j) GLM_LOG_B##j
#define GLM_POWLOG_TAB(i, f) __ldg(&glm_powlog_tab[4 * (i) + (f)])
#define GLM_LOG_TAB(i, f) __ldg(&glm_log_tab[2 * (i) + (f)])
#define GLM_EXP_TAB(i) __ldg(&glm_exp_tab[(i)])
#include "glibc_math_core.h"

// out of line: one copy per kernel instead of one per call site (17 pow sites), see land_common.cuh HDIV
static __device__ __noinline__ double hsp_pow(double x, double y) {
    double r;
    if (glm_pow_main(x, y, &r)) return r;
    return pow(x, y);
}
static __device__ __noinline__ double hsp_exp(double x) {
    double r;
    if (glm_exp_main(x, &r)) return r;
    return exp(x);
}
static __device__ __noinline__ double hsp_exp2(double x) {
    double r;
    if (glm_exp2_main(x, &r)) return r;
    return exp2(x);
}
static __device__ __noinline__ double hsp_log(double x) {
    double r;
    if (glm_log_main(x, &r)) return r;
    return log(x);
}
#endif
