// land_common.cuh -- launch record, per-thread accessors and the forcing pipeline shared by the PERLND and
// IMPLND kernels.  sm_100a only; compiled with -fmad=false (the reference has no FMA contraction).
//
// Work decomposition: one thread = one land segment, one warp = one TILE of 32 segments.  Warps are fully
// independent (no block-level barrier anywhere), so every buffer is laid out tile-major and a warp's data is one
// contiguous span whose rows sit at COMPILE-TIME offsets from a single per-thread pointer:
//
//   segblk [tile][SB_ROWS][32]          8-byte cells: row 0 = 64-bit option word, then HSP2G_NPAR parameters,
//                                       HSP2G_NSTATE carried states (read at chunk start, written at chunk end),
//                                       HSP2G_NERR int64 error counters
//   mon    [tile][slot*12 + month][32]  monthly tables that at least one segment uses
//   forc   [tile][t][row][32]           time-major forcings of the chunk: interval t of a tile is nf*256
//                                       contiguous bytes -> ONE 1-D TMA bulk copy (cp.async.bulk) per warp-interval
//   out    [tile][t][row][32]           time-major saved series: a warp stores 256 contiguous bytes per row
//   cal    [steps+1]                    shared per-interval calendar record (warp-uniform loads)
//
// The last tile is padded to 32 lanes with copies of the last real segment (hsp2g_api.cu), so padded lanes run
// the same control flow as a real one and nothing in the kernels is masked.
#pragma once
#include "cuda_rt.h"
#include <stdint.h>

#include "hsp2g_ids.h"

#define LAND_TILE 32
#define LAND_BLOCK 128
#define LAND_WARPS (LAND_BLOCK / LAND_TILE)
#define PIPE_STAGES 3

#define SB_FLAGS 0
#define SB_PAR 1
#define SB_STATE (SB_PAR + HSP2G_NPAR)
#define SB_ERR (SB_STATE + HSP2G_NSTATE)
#define SB_ROWS (SB_ERR + HSP2G_NERR)

struct __align__(32) CalStep {
    double xm;       // (day - first of month) in the calendar's time unit
    double dxm;      // month length in the same unit
    double lapse;    // ATEMP dry lapse rate of this interval (ATEMP.py:19)
    uint8_t flags;   // HSP2G_CAL_*
    uint8_t m0, m1;  // month indices of the bracketing table entries
    uint8_t pad[5];
};

struct LandLaunch {
    double *segblk;
    const double *mon;
    const double *forc;
    double *out;
    const CalStep *cal;  // already offset to the first interval of this chunk
    int ntiles;
    int nsteps;
    int first_chunk;  // 1 when this chunk starts at interval 0 of the run
    int nf, ns;       // forcing rows / saved rows per interval
    int mon_rows;     // 12 * number of monthly slots (row count of one tile's table block)
    unsigned act;
    unsigned opts;
    double delt, delt60;
    short frow[HSP2G_NFORC];    // forcing id -> row in forc, or -1
    double ffill[HSP2G_NFORC];  // value read when the id is absent
    short srow[HSP2G_NOUT];     // output id -> row in out, or -1 (not saved)
    short mslot[HSP2G_NMON];    // monthly table id -> slot in mon, or -1
};

// Python semantics: max(a,b) = b if b>a else a ; min(a,b) = b if b<a else a  
__device__ __forceinline__ double py_max(double a, double b) { return b > a ? b : a; }
__device__ __forceinline__ double py_min(double a, double b) { return b < a ? b : a; }

#ifdef HSP2G_HOST_EMU
__device__ __forceinline__ int hsp_popc(unsigned long long x) { return __builtin_popcountll(x); }
#else
__device__ __forceinline__ int hsp_popc(unsigned long long x) { return __popcll(x); }  // folds on constants
#endif

// ---- I/O policy: which forcing rows exist and which outputs are saved --------------------------------------
// DynIO reads the row tables of the launch record at run time (any layout).  StaticIO<forcing mask, save mask>
// fixes both sets at compile time with rows in ascending id order: absent outputs cost nothing (their stores and
// the arithmetic that only feeds them disappear) and every present row is an immediate offset.
struct DynIO {
    static constexpr bool STATIC = false;
    static constexpr unsigned long long FMASK = 0, S0 = 0, S1 = 0;
    static constexpr int NF = 0, NS = 0;
};
constexpr int ce_popc(unsigned long long x) { return x ? (int)(x & 1ull) + ce_popc(x >> 1) : 0; }
template <unsigned long long FM, unsigned long long SM0, unsigned long long SM1>
struct StaticIO {
    static constexpr bool STATIC = true;
    static constexpr unsigned long long FMASK = FM, S0 = SM0, S1 = SM1;
    static constexpr int NF = ce_popc(FM), NS = ce_popc(SM0) + ce_popc(SM1);
};
static_assert(HSP2G_NFORC <= 64 && HSP2G_NOUT <= 128, "StaticIO masks are 64 + 128 bits");

#ifndef HSP2G_HOST_EMU
// ---- forcing pipeline: a PIPE_STAGES-deep ring per warp in shared memory, filled by the TMA engine.  Lane 0 arms
// the stage's mbarrier with the byte count and issues one cp.async.bulk for the whole interval (nf rows x 256 B,
// contiguous in HBM thanks to the tile-major layout); all lanes wait on the barrier phase.  No LSU traffic, no
// register staging, no block barrier.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "HSP_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra HSP_DONE;\n"
        "bra HSP_WAIT;\n"
        "HSP_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}

struct Pipe {
    double *ring;       // this warp's ring (generic pointer into shared memory)
    unsigned ring_s;    // its shared-space address
    unsigned bar_s;     // shared-space address of this warp's first mbarrier
    const double *src;  // forc + tile * nsteps * nf * 32
    int stage_elems;    // nf * 32
    int nsteps;
    int st;             // stage of the current interval
    unsigned ph;        // its phase parity
    int lane;

    __device__ __forceinline__ void init(double *smem, int wib, int lane_, const double *src_, int nf, int nsteps_) {
        stage_elems = nf * LAND_TILE;
        ring = smem + (size_t)wib * PIPE_STAGES * stage_elems;
        ring_s = smem_u32(ring);
        bar_s = smem_u32(smem + (size_t)LAND_WARPS * PIPE_STAGES * stage_elems) + wib * PIPE_STAGES * 8;
        src = src_;
        nsteps = nf > 0 ? nsteps_ : 0;
        st = 0;
        ph = 0;
        lane = lane_;
        if (lane == 0) {
            for (int s = 0; s < PIPE_STAGES; ++s) mbar_init(bar_s + s * 8, 1);
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            for (int s = 0; s < PIPE_STAGES; ++s) issue(s, s);
        }
        __syncwarp();
    }
    __device__ __forceinline__ void issue(int t, int s) {  // lane 0 only
        if (t < nsteps) {
            const unsigned bytes = (unsigned)stage_elems * 8u;
            mbar_expect_tx(bar_s + s * 8, bytes);
            bulk_g2s(ring_s + (unsigned)s * bytes, src + (size_t)t * stage_elems, bytes, bar_s + s * 8);
        }
    }
    // wait for interval t's forcings; returns this lane's column of the stage
    __device__ __forceinline__ const double *acquire(int t) {
        if (nsteps > 0) mbar_wait(bar_s + st * 8, ph);
        return ring + st * stage_elems + lane;
    }
    // all lanes are done with interval t: refill its stage with interval t + PIPE_STAGES
    __device__ __forceinline__ void release(int t) {
        __syncwarp();
        if (lane == 0) issue(t + PIPE_STAGES, st);
        if (++st == PIPE_STAGES) {
            st = 0;
            ph ^= 1u;
        }
    }
};
#define HSP_SYNCWARP() __syncwarp()

#else
// test double: same interface, synchronous copies (no TMA on the host); one "thread" at a time
struct Pipe {
    double *ring;
    const double *src;
    int stage_elems, nsteps, lane;
    void init(double *smem, int, int lane_, const double *src_, int nf, int nsteps_) {
        ring = smem;
        src = src_;
        stage_elems = nf * LAND_TILE;
        nsteps = nsteps_;
        lane = lane_;
    }
    const double *acquire(int t) {
        for (int i = lane; i < stage_elems; i += LAND_TILE) ring[i] = src[(size_t)t * stage_elems + i];
        return ring + lane;
    }
    void release(int) {}
};
#define HSP_SYNCWARP()
#endif

// per-thread view of the batch
template <class IO_>
struct Seg {
    typedef IO_ IO;
    const LandLaunch &L;
    double *sb;            // this lane's column of the tile's segment block
    const double *monp;    // this lane's column of the tile's monthly tables
    double *outp;          // this lane's column of out[tile][t]
    const double *stage;   // this lane's column of the current forcing stage (shared memory)
    const CalStep *calp;   // calendar record of the current interval (warp-uniform)
    unsigned long long fl;
    unsigned calfl;
    int out_step;          // elements between consecutive intervals of out

    __device__ __forceinline__ Seg(const LandLaunch &l, int tile, int lane) : L(l) {
        sb = l.segblk + ((size_t)tile * SB_ROWS * LAND_TILE + lane);
        monp = l.mon + ((size_t)tile * l.mon_rows * LAND_TILE + lane);
        out_step = (IO::STATIC ? IO::NS : l.ns) * LAND_TILE;
        outp = l.out + ((size_t)tile * l.nsteps * out_step + lane);
        calp = l.cal;
        fl = *reinterpret_cast<const unsigned long long *>(sb + SB_FLAGS * LAND_TILE);
        stage = nullptr;
        calfl = 0;
    }
    __device__ __forceinline__ void begin_step(const double *stage_) {
        stage = stage_;
        calfl = calp->flags;
    }
    __device__ __forceinline__ void end_step() {
        outp += out_step;
        ++calp;
    }
    __device__ __forceinline__ bool flag(int bit) const { return (fl >> bit) & 1ull; }
    __device__ __forceinline__ double par(int id) const { return __ldg(sb + (SB_PAR + id) * LAND_TILE); }
    __device__ __forceinline__ bool has_forcing(int id) const {
        if (IO::STATIC) return (IO::FMASK >> id) & 1ull;
        return L.frow[id] >= 0;
    }
    __device__ __forceinline__ double forcing(int id) const {
        if (IO::STATIC) {
            if ((IO::FMASK >> id) & 1ull) return stage[hsp_popc(IO::FMASK & ((1ull << id) - 1ull)) * LAND_TILE];
            return L.ffill[id];
        }
        const int r = L.frow[id];
        return r >= 0 ? stage[r * LAND_TILE] : L.ffill[id];
    }
    // monthly table value for the current day: numpy's slope*(x-x0)+y0 (calendar.py dayval), no FMA
    __device__ __forceinline__ double monthly(int mid) const {
        const double *tab = monp + (size_t)L.mslot[mid] * (12 * LAND_TILE);
        const double y0 = __ldg(tab + (int)calp->m0 * LAND_TILE);
        const double xm = calp->xm;
        if (xm == 0.0) return y0;
        const double y1 = __ldg(tab + (int)calp->m1 * LAND_TILE);
        const double slope = (y1 - y0) / calp->dxm;
        return slope * xm + y0;
    }
    // initm(): the table where the segment's flag is on, else the constant parameter
    __device__ __forceinline__ double daily(int flagbit, int mid, int pid) const {
        return flag(flagbit) ? monthly(mid) : par(pid);
    }
    __device__ __forceinline__ bool saved(int oid) const {
        if (IO::STATIC) return ((oid < 64 ? IO::S0 : IO::S1) >> (oid & 63)) & 1ull;
        return L.srow[oid] >= 0;
    }
    __device__ __forceinline__ void save(int oid, double v) const {
        if (IO::STATIC) {
            if (((oid < 64 ? IO::S0 : IO::S1) >> (oid & 63)) & 1ull) {
                const int r = (oid < 64 ? 0 : hsp_popc(IO::S0)) +
                              hsp_popc((oid < 64 ? IO::S0 : IO::S1) & ((1ull << (oid & 63)) - 1ull));
                __stcs(outp + r * LAND_TILE, v);
            }
        } else {
            const int r = L.srow[oid];
            if (r >= 0) __stcs(outp + r * LAND_TILE, v);
        }
    }
    __device__ __forceinline__ double st(int sid) const { return sb[(SB_STATE + sid) * LAND_TILE]; }
    __device__ __forceinline__ void st(int sid, double v) const { sb[(SB_STATE + sid) * LAND_TILE] = v; }
    __device__ __forceinline__ void count(int eid, int n) const {
        if (n) reinterpret_cast<long long *>(sb)[(SB_ERR + eid) * LAND_TILE] += n;
    }
};
