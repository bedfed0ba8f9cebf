// land_common.cuh -- launch record, per-thread accessors and the forcing pipeline shared by the PERLND and
// IMPLND kernels.  sm_100a only; compiled with -fmad=false (the reference has no FMA contraction).
//
// Data layout in HBM (n_pad = segment count rounded up to 128, every row 1 KiB-aligned):
//   par   [HSP2G_NPAR][n_pad]            fp64 scalar parameters, structure of arrays
//   mon   [slot][12][n_pad]              monthly tables that at least one segment uses
//   flags [n_pad]                        64-bit option word per segment
//   state [HSP2G_NSTATE][n_pad]          carried state (named + hidden), read at chunk start, written at end
//   err   [HSP2G_NERR][n_pad]            int64 error counters, accumulated
//   forc  [t][nf][n_pad]                 time-major forcings: a warp reads 256 contiguous bytes per (t, row)
//   out   [t][ns][n_pad]                 time-major saved series: a warp writes 256 contiguous bytes
//   cal   [steps+1]                      shared per-step calendar record (broadcast loads)
#pragma once
#include "cuda_rt.h"
#include <stdint.h>

#include "hsp2g_ids.h"

struct __align__(32) CalStep {
    double xm;       // (day - first of month) in the calendar's time unit
    double dxm;      // month length in the same unit
    double lapse;    // ATEMP dry lapse rate of this interval (ATEMP.py:19)
    uint8_t flags;   // HSP2G_CAL_*
    uint8_t m0, m1;  // month indices of the bracketing table entries
    uint8_t pad[5];
};

struct LandLaunch {
    const double *par;
    const double *mon;
    const unsigned long long *flags;
    double *state;
    long long *err;
    const double *forc;
    double *out;
    const CalStep *cal;  // already offset to the first step of this chunk
    long long n_pad;
    int n_seg;
    int nsteps;
    int first_chunk;  // 1 when this chunk starts at step 0 of the run
    int nf, ns;
    unsigned act;
    unsigned opts;
    double delt, delt60;
    short frow[HSP2G_NFORC];    // forcing id -> row in forc, or -1
    double ffill[HSP2G_NFORC];  // value read when the id is absent
    short srow[HSP2G_NOUT];     // output id -> row in out, or -1 (not saved)
    short mslot[HSP2G_NMON];    // monthly table id -> slot in mon, or -1
};

#define LAND_BLOCK 128
#define PIPE_STAGES 4

// Python semantics: max(a,b) = b if b>a else a ; min(a,b) = b if b<a else a  
__device__ __forceinline__ double py_max(double a, double b) { return b > a ? b : a; }
__device__ __forceinline__ double py_min(double a, double b) { return b < a ? b : a; }

#ifndef HSP2G_HOST_EMU
// ---- forcing pipeline: each thread owns a private PIPE_STAGES-deep ring of its own forcing values in
// shared memory, filled with 8-byte cp.async (LDGSTS) so no barrier and no register staging is needed.
// smem layout [stage][row][thread] keeps every access conflict-free.
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

struct Pipe {
    double *ring;        // this thread's column: ring[(stage*nf + row)*LAND_BLOCK]
    const double *src;   // forc + seg
    long long step_stride;  // nf * n_pad
    long long n_pad;
    int nf;
    int nsteps;

    __device__ __forceinline__ void issue(int t) {
        if (t < nsteps) {
            double *dst = ring + (size_t)(t % PIPE_STAGES) * nf * LAND_BLOCK;
            const double *s = src + (long long)t * step_stride;
            for (int r = 0; r < nf; ++r) cp_async8(dst + r * LAND_BLOCK, s + (long long)r * n_pad);
        }
        cp_async_commit();  // empty groups keep the wait arithmetic uniform at the tail
    }
    __device__ __forceinline__ void prologue() {
        for (int t = 0; t < PIPE_STAGES - 1; ++t) issue(t);
    }
    // make step t's values visible and start the copy PIPE_STAGES-1 steps ahead
    __device__ __forceinline__ const double *acquire(int t) {
        cp_async_wait<PIPE_STAGES - 2>();
        return ring + (size_t)(t % PIPE_STAGES) * nf * LAND_BLOCK;
    }
    __device__ __forceinline__ void release(int t) { issue(t + PIPE_STAGES - 1); }
};

#else
// test double: same interface, synchronous copies (no cp.async on the host)
struct Pipe {
    double *ring;
    const double *src;
    long long step_stride;
    long long n_pad;
    int nf;
    int nsteps;
    void prologue() {}
    const double *acquire(int t) {
        for (int r = 0; r < nf; ++r) ring[r * LAND_BLOCK] = src[(long long)t * step_stride + (long long)r * n_pad];
        return ring;
    }
    void release(int) {}
};
template <int N>
inline void cp_async_wait() {}
#endif

// per-thread view of the batch
struct Seg {
    const LandLaunch &L;
    long long seg;
    unsigned long long fl;
    const double *stage;  // current forcing stage (thread column)
    CalStep cal;
    int t;

    __device__ __forceinline__ Seg(const LandLaunch &l, long long s) : L(l), seg(s) { fl = l.flags[s]; }
    __device__ __forceinline__ bool flag(int bit) const { return (fl >> bit) & 1ull; }
    __device__ __forceinline__ double par(int id) const { return __ldg(L.par + (long long)id * L.n_pad + seg); }
    __device__ __forceinline__ double forcing(int id) const {
        const int r = L.frow[id];
        return r >= 0 ? stage[r * LAND_BLOCK] : L.ffill[id];
    }
    __device__ __forceinline__ bool has_forcing(int id) const { return L.frow[id] >= 0; }
    // monthly table value for the current day: numpy's slope*(x-x0)+y0 (calendar.py dayval), no FMA
    __device__ __forceinline__ double monthly(int mid) const {
        const double *tab = L.mon + ((long long)L.mslot[mid] * 12) * L.n_pad + seg;
        const double y0 = __ldg(tab + (long long)cal.m0 * L.n_pad);
        if (cal.xm == 0.0) return y0;
        const double y1 = __ldg(tab + (long long)cal.m1 * L.n_pad);
        const double slope = (y1 - y0) / cal.dxm;
        return slope * cal.xm + y0;
    }
    // initm(): the table where the segment's flag is on, else the constant parameter
    __device__ __forceinline__ double daily(int flagbit, int mid, int pid) const {
        return flag(flagbit) ? monthly(mid) : par(pid);
    }
    __device__ __forceinline__ void save(int oid, double v) const {
        const int r = L.srow[oid];
        if (r >= 0) __stcs(L.out + ((long long)t * L.ns + r) * L.n_pad + seg, v);
    }
    __device__ __forceinline__ double st(int sid) const { return L.state[(long long)sid * L.n_pad + seg]; }
    __device__ __forceinline__ void st(int sid, double v) const { L.state[(long long)sid * L.n_pad + seg] = v; }
    __device__ __forceinline__ void count(int eid, int n) const {
        if (n) L.err[(long long)eid * L.n_pad + seg] += n;
    }
};
