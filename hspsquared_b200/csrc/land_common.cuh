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
//   forc   TILED  [tile][t][row][32]    time-major forcings of the chunk: interval t of a tile is nf*256
//                                       contiguous bytes -> ONE 1-D TMA bulk copy (cp.async.bulk) per warp-interval
//          PLAIN  [t][row][ld]          the reference's own orientation ([t][segment] per series, ld >= 32*ntiles): interval t
//                                       of a tile is a box of nf rows x 32 segments -> ONE tensor-map TMA copy
//                                       (cp.async.bulk.tensor.3d) per warp-interval; no re-tiling on either side of PCIe
//   out    TILED  [tile][t][row][32]    time-major saved series: a warp stores 256 contiguous bytes per row
//          PLAIN  [t][row][ld]          ... the same 256-byte stores, rows ld elements apart
//          forcings may be float32 (widened on the way out of shared memory), saved series float32 or float64
//   cal    [steps+1]                    shared per-interval calendar record (warp-uniform loads)
//
// The last tile is padded to 32 lanes with copies of the last real segment (hsp2g_api.cu), so padded lanes run
// the same control flow as a real one and nothing in the kernels is masked.
#pragma once
#include "cuda_rt.h"
#include <stdint.h>

#include "glibc_math.cuh"
#include "hsp2g_ids.h"

#define LAND_TILE 32
#ifndef LAND_BLOCK
#define LAND_BLOCK 64  // two independent warps per block: warps never synchronise; pairs halve the 1 KB per-block smem reservation
#endif
#define LAND_WARPS (LAND_BLOCK / LAND_TILE)
#ifndef PIPE_STAGES
#define PIPE_STAGES 2  // one interval of prefetch covers HBM latency (an interval takes a warp microseconds)
#endif

#ifdef HSP2G_HOST_EMU
#define LAND_SMEM_DECL double *smem = ring
#define HSP_MAXNREG(n)
#else
#define LAND_SMEM_DECL extern __shared__ __align__(128) double smem[]
#define HSP_MAXNREG(n) __maxnreg__(n)
#endif

#define SB_FLAGS 0
#define SB_PAR 1
#define SB_STATE (SB_PAR + HSP2G_NPAR)
#define SB_ERR (SB_STATE + HSP2G_NSTATE)
#define SB_ROWS (SB_ERR + HSP2G_NERR)

// Parameters read on (nearly) every interval are staged once per launch into this warp's shared memory: an L1
// miss on a parameter row costs an L2 round trip in the middle of a serial fp64 dependency chain (profiles/r1b).
#define PBIT(n) (1ull << P_##n)
// Per warp (headline kernel): 2 stages x 6 forcing rows + 23 hot parameters + 19 cold states = 54 rows of 256 B
// = 13.5 KB; 8 two-warp blocks (+1 KB system reservation each) fit the 228 KB of an SM = 16 resident warps at <= 128
// registers (RDCSN, read only while it snows, stays in global memory: a 55th row would cost two warps).
#define HOT_SNOW                                                                                                   \
    (PBIT(TSNOW) | PBIT(SNOWCF) | PBIT(COVIND) | PBIT(SNOEVP_K) | PBIT(CCFACT_K) | PBIT(MELEV_FAC) |  \
     PBIT(SHADE) | PBIT(MWATER) | PBIT(MGMELT_IVL))
#define HOT_PWATER                                                                                                 \
    (PBIT(INFILT_IVL) | PBIT(KVARY) | PBIT(LZSN) | PBIT(FOREST) | PBIT(PETMAX) | PBIT(PETMIN) | PBIT(FZG) |         \
     PBIT(FZGL) | PBIT(DEEPFR) | PBIT(KGW) | PBIT(BASETP) | PBIT(AGWETP) | PBIT(INFEXP) | PBIT(INFILD))
#define HOT_IWATER (PBIT(PETMAX) | PBIT(PETMIN))
// full PERLND chain: ATEMP / SEDMNT / PWTGAS constants read on every (wet) interval
#define HOT_CHAIN                                                                                                  \
    (PBIT(ELDAT) | PBIT(SMPF) | PBIT(KRER) | PBIT(JRER) | PBIT(KSER) | PBIT(JSER) | PBIT(KGER) | PBIT(JGER) |       \
     PBIT(ELEVGC) | PBIT(SLIFAC) | PBIT(ILIFAC) | PBIT(ALIFAC))
static_assert(P_RETSC < 64, "hot-parameter masks are 64 bits: hot ids must be below 64 (later ids are never hot)");

struct __align__(32) CalStep {
    double xm;       // (day - first of month) in the calendar's time unit
    double dxm;      // month length in the same unit
    double lapse;    // ATEMP dry lapse rate of this interval (ATEMP.py:19)
    uint8_t flags;   // HSP2G_CAL_*
    uint8_t m0, m1;  // month indices of the bracketing table entries
    uint8_t pad[5];
};

// cache policy of the saved-series stores: streaming (evict-first) by default; HSP_STORE_POLICY is a measurement switch
// (1 = ordinary write-back, 2 = .cg, 3 = write-through), profiles/README.md
#if !defined(HSP_STORE_POLICY) || HSP_STORE_POLICY == 0 || defined(HSP2G_HOST_EMU)
#define HSP_STORE(p, v) __stcs((p), (v))
#elif HSP_STORE_POLICY == 1
#define HSP_STORE(p, v) (*(p) = (v))
#elif HSP_STORE_POLICY == 2
#define HSP_STORE(p, v) __stcg((p), (v))
#else
#define HSP_STORE(p, v) __stwt((p), (v))
#endif

// Linked forcing rows (hsp2g_run_device_linked): row r of interval t of this batch's tile q is the 256 bytes at
// base + t * t_stride + (q % src_tiles) * tile_stride -- a row of ANOTHER batch's device buffer, read in place.
#define LAND_MAX_LINKED 16
struct RowSource {
    const unsigned char *base;
    long long t_stride, tile_stride;  // bytes
};

struct LandLaunch {
    double *segblk;
    const double *mon;
    const double *forc;
    double *out;
    const CalStep *cal;  // already offset to the first interval of this chunk
    // shared-met mode (hsp2g_set_shared_forcing): forc is unused; interval t, row r of a segment is
    // met[(t*nf + r)*n_stations + station[seg]] * mfac[tile][r][lane]   (get_timeseries' series * MFACTOR)
    const double *met;   // already offset to the first interval of this chunk; nullptr = per-segment forcings
    const int *station;  // [tile][32]
    const double *mfac;  // [tile][nf][32]
    int n_stations;
    int met_table;       // shared-met: 1 = the interval's whole station table is fetched by one bulk copy (met_table_mode)
    int ntiles;
    int nsteps;
    int sub;          // intervals per task; a task = (tile, sub-chunk) (see Sched)
    int nsub;         // sub-chunks per tile in this launch = ceil(nsteps / sub)
    unsigned long long *task_counter;  // global ticket counter, monotonically increasing over the batch's life
    unsigned long long task_base;      // its value when this launch starts
    unsigned *progress;                // [ntiles] sub-chunks completed by each tile since the batch was created
    unsigned prog_base;                // their common value when this launch starts
    int first_chunk;  // 1 when this chunk starts at interval 0 of the run
    int nf, ns;       // forcing rows / saved rows per interval
    int out_f32;      // saved series are float32 instead of float64
    int plain;        // series layout: 0 = TILED [tile][t][row][32], 1 = PLAIN [t][row][ld]
    int forc_f32;     // forcing series are float32 (exactly representable inputs, e.g. WDM / HDF5 float32 stores)
    int tmap_ok;      // PLAIN: `tmap` describes forc (else the rows of an interval are fetched by nf 1-D bulk copies)
    long long forc_ld, out_ld;  // PLAIN: elements between consecutive rows of forc / out
    int mon_rows;     // 12 * number of monthly slots (row count of one tile's table block)
    unsigned act;
    unsigned opts;
    int flags_uniform;              // 1: every segment of the batch carries the option word flags_word
    unsigned long long flags_word;
    unsigned long long flags_and, flags_or;  // AND / OR of all segments' option words: the bits they agree on are ~(and ^ or)
    double delt, delt60;
    short frow[HSP2G_NFORC];    // forcing id -> row in forc, or -1
    double ffill[HSP2G_NFORC];  // value read when the id is absent
    short srow[HSP2G_NOUT];     // output id -> row in out, or -1 (not saved)
    short mslot[HSP2G_NMON];    // monthly table id -> slot in mon, or -1
    int chained;      // 1: the previous launch of this batch may still be running (hsp2g_set_launch_overlap): sub-chunk 0 of a tile
                      // waits for progress[tile] like every other sub-chunk does
    int linked;       // 1: the forcing rows come from rowsrc (forc, forc_ld, tmap are unused)
    int src_tiles;    // linked: tiles of the source batch; this batch's tile q reads source tile q % src_tiles
    RowSource rowsrc[LAND_MAX_LINKED];
    // PLAIN: CUtensorMap of forc as a 3-D tensor (segment, row, interval), box = 32 x nf x 1 (hsp2g_api.cu make_tmap)
    alignas(64) unsigned long long tmap[16];
};

// A carried scalar that lives in this lane's column of the warp's shared-memory scratch instead of a register.
// Used for state that changes slowly (once a day) or that an hourly run only needs again at the end of the chunk:
// ~23 doubles = 46 registers per thread that would otherwise force either spills or fewer resident warps.
// (p may also point at the value's own row of the segment block in global memory, for the rarely touched ones.)
struct ColdD {
    double *p;
    __device__ __forceinline__ operator double() const { return *p; }
    __device__ __forceinline__ ColdD &operator=(double v) {
        *p = v;
        return *this;
    }
    __device__ __forceinline__ ColdD &operator=(const ColdD &o) {
        *p = *o.p;
        return *this;
    }
    __device__ __forceinline__ ColdD &operator+=(double v) { return *this = *p + v; }
    __device__ __forceinline__ ColdD &operator-=(double v) { return *this = *p - v; }
    __device__ __forceinline__ ColdD &operator*=(double v) { return *this = *p * v; }
};
// Same idea for values only wet intervals touch, left in their state row in global memory.  L2-only accesses: a
// tile's sub-chunks may run on different SMs within one launch, whose L1s are not coherent.
#ifdef HSP2G_HOST_EMU
typedef ColdD ColdG;
#else
struct ColdG {
    double *p;
    __device__ __forceinline__ operator double() const { return __ldcg(p); }
    __device__ __forceinline__ ColdG &operator=(double v) {
        __stcg(p, v);
        return *this;
    }
};
#endif
// A carried scalar that an hourly kernel recomputes before every use: the register copy `v` is then dead across
// the time loop (no register, no shared memory), and each assignment writes the value through to its state row in
// global memory so the next task / chunk finds it.  Where some use is NOT dominated by an assignment (an output saved
// every interval, say) the compiler simply keeps `v` live, so the type is safe for any configuration.
#ifdef HSP2G_HOST_EMU
struct ColdW {
    double v, *p;
    operator double() const { return v; }
    ColdW &operator=(double x) {
        v = x;
        *p = x;
        return *this;
    }
};
#else
struct ColdW {
    double v, *p;
    __device__ __forceinline__ operator double() const { return v; }
    __device__ __forceinline__ ColdW &operator=(double x) {
        v = x;
        __stcg(p, x);
        return *this;
    }
};
#endif
#define NCOLD_SNOW 12
#define NCOLD_SNOW2 7
#define NCOLD_PWATER 12
#define NCOLD_PSTEMP 6
#define NCOLD_PWTGAS 4
#define NCOLD_CHAIN (NCOLD_PSTEMP + NCOLD_PWTGAS)

// One out-of-line copy of the IEEE fp64 division sequence (MUFU.RCP64H + Newton + residual, ~20 SASS
// instructions inline at each of ~50 sites = a quarter of the kernel).  a / b, correctly rounded, same bits.
static __device__ __noinline__ double hsp_div(double a, double b) { return a / b; }
#ifdef HSP_INLINE_DIV
#define HDIV(a, b) ((a) / (b))
#else
#define HDIV(a, b) hsp_div((a), (b))
#endif

// a / C for the constants whose odd part is a small integer: three operations, same bits (const_div.h has the proof and the pin)
#ifdef HSP2G_HOST_EMU
static inline unsigned hsp_hi32(double x) {
    unsigned long long u;
    memcpy(&u, &x, 8);
    return (unsigned)(u >> 32);
}
#define HSP_CDIV_DECL static inline
#define HSP_CDIV_FMA(a, b, c) __builtin_fma((a), (b), (c))
#define HSP_CDIV_HI(x) hsp_hi32(x)
#else
#define HSP_CDIV_DECL static __device__ __forceinline__
#define HSP_CDIV_FMA(a, b, c) __fma_rn((a), (b), (c))
#define HSP_CDIV_HI(x) __double2hiint(x)
#endif
#define HSP_CDIV_SLOW(a, c) hsp_div((a), (c))
#include "const_div.h"

// Python semantics: max(a,b) = b if b>a else a ; min(a,b) = b if b<a else a  
__device__ __forceinline__ double py_max(double a, double b) { return b > a ? b : a; }
__device__ __forceinline__ double py_min(double a, double b) { return b < a ? b : a; }

#ifdef HSP2G_HOST_EMU
__device__ __forceinline__ int hsp_popc(unsigned long long x) { return __builtin_popcountll(x); }
#else
__device__ __forceinline__ int hsp_popc(unsigned long long x) { return __popcll(x); }  // folds on constants
#endif
template <class IO>
__device__ __forceinline__ int io_row(int oid) {
    return (oid < 64 ? 0 : hsp_popc(IO::S0)) + (oid < 128 ? 0 : hsp_popc(IO::S1)) +
           hsp_popc((oid < 64 ? IO::S0 : oid < 128 ? IO::S1 : IO::S2) & ((1ull << (oid & 63)) - 1ull));
}

// ---- I/O policy: which forcing rows exist and which outputs are saved --------------------------------------
// DynIO reads the row tables of the launch record at run time (any layout).  StaticIO<forcing mask, save mask>
// fixes both sets at compile time with rows in ascending id order: absent outputs cost nothing (their stores and
// the arithmetic that only feeds them disappear) and every present row is an immediate offset.
// F32: saved series are stored as float32 (round to nearest even), which is what the reference hands its IOManager
// (save_timeseries casts the frame with astype(np.float32), utilities.py:313): half the bytes to HBM and over PCIe.
struct DynIO {
    static constexpr bool STATIC = false, F32 = false;  // DynIO reads L.out_f32 at run time
    static constexpr unsigned long long FMASK = 0, S0 = 0, S1 = 0, S2 = 0;
    static constexpr int NF = 0, NS = 0;
};
__host__ __device__ constexpr int ce_popc(unsigned long long x) { return x ? (int)(x & 1ull) + ce_popc(x >> 1) : 0; }
// forcing ids < 64 in FM; output ids 0..63 in SM0, 64..127 in SM1, 128..191 in SM2
template <unsigned long long FM, unsigned long long SM0, unsigned long long SM1, bool F32_ = false, unsigned long long SM2 = 0>
struct StaticIO {
    static constexpr bool STATIC = true, F32 = F32_;
    static constexpr unsigned long long FMASK = FM, S0 = SM0, S1 = SM1, S2 = SM2;
    static constexpr int NF = ce_popc(FM), NS = ce_popc(SM0) + ce_popc(SM1) + ce_popc(SM2);
};
static_assert(HSP2G_NFORC <= 64 && HSP2G_NOUT <= 192, "StaticIO masks are 64 + 192 bits");
// compile-time SAVE set: is output `oid` saved; its row (rows are in ascending id order) is io_row<IO>() above, written with
// the device population count, which folds on constants (ce_popc is a host constexpr and must not be reached at run time).
template <class IO>
__host__ __device__ constexpr bool io_saved(int oid) {
    return ((oid < 64 ? IO::S0 : oid < 128 ? IO::S1 : IO::S2) >> (oid & 63)) & 1ull;
}
#ifdef HSP2G_HOST_EMU
#define HSP_PIN(p)
#define HSP_PIN32(v)
#else
// stop ptxas from re-deriving a loop-invariant address (or the lane index itself) from threadIdx / blockIdx / the shared
// window at every use: under register pressure it rematerialises them (an S2R and a few integer operations each time;
// ~8 % of all issued instructions in profiles/r1b, 48 S2R in a kernel built without the pins in round 2)
#define HSP_PIN(p) asm volatile("" : "+l"(p))
#define HSP_PIN32(v) asm volatile("" : "+r"(v))
#endif

// Shared-met mode with few stations: the whole station table of an interval ([nf][n_stations] doubles, contiguous in the met
// array) is fetched by ONE bulk copy per warp-interval and every lane picks its station's column out of shared memory, instead
// of nf 8-byte gathers per lane (cp.async).  Needs the table to fit a stage (n_stations <= 32) and to be a multiple of 16 bytes.
// The host decides (hsp2g_api.cu launch(): L.met_table); kernels built for one batch have the answer at compile time (StaticLay).
__host__ __device__ inline bool met_table_fits(int n_stations, int nf) {
    return n_stations > 0 && n_stations <= LAND_TILE && ((nf * n_stations) & 1) == 0;
}

// ---- series layout policy: where interval t, row r of a tile lives (see the header of this file) and the element type
// of the forcings.  StaticLay fixes both at compile time (kernels built for one batch, hspsquared_b200/jit.py);
// DynLay reads them from the launch record (the general kernels of the library).
struct DynLay {
    static constexpr bool STATIC = false, PLAIN = false, F32IN = false, LINKED = false;
    static constexpr int METTBL = -1;
};
// METTBL: shared-met table mode 0 / 1, -1 = read L.met_table; LINKED: forcing rows read in place from another batch's buffers
template <bool PLAIN_, bool F32IN_, int METTBL_ = -1, bool LINKED_ = false>
struct StaticLay {
    static constexpr bool STATIC = true, PLAIN = PLAIN_, F32IN = F32IN_, LINKED = LINKED_;
    static constexpr int METTBL = METTBL_;
};
typedef StaticLay<false, false> TiledLay;
template <class LAY>
__device__ __forceinline__ bool lay_plain(const LandLaunch &L) {
    return LAY::STATIC ? LAY::PLAIN : (L.plain != 0);
}
template <class LAY>
__device__ __forceinline__ bool lay_f32in(const LandLaunch &L) {
    return LAY::STATIC ? LAY::F32IN : (L.forc_f32 != 0);
}
template <class LAY>
__device__ __forceinline__ bool lay_linked(const LandLaunch &L) {
    return LAY::STATIC ? LAY::LINKED : (L.linked != 0);
}
template <class LAY>
__device__ __forceinline__ bool lay_mettbl(const LandLaunch &L) {
    return LAY::METTBL >= 0 ? LAY::METTBL != 0 : (L.met_table != 0);
}

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
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"  // %2: suspend-time hint, the warp sleeps
        "@p bra HSP_DONE;\n"                                            // in hardware instead of polling
        "bra HSP_WAIT;\n"
        "HSP_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity), "r"(0x989680u)
        : "memory");
}

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem) : "memory");
}

// 3-D tensor-map copy: the box at (segment c0, row 0, interval c2) of the PLAIN forcing tensor -> shared memory
__device__ __forceinline__ void tensor_g2s(unsigned dst, const void *tmap, int c0, int c2, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(dst),
        "l"(tmap), "r"(c0), "r"(0), "r"(c2), "r"(bar)
        : "memory");
}

// SHARED = false: per-segment forcings, one TMA copy per warp-interval: a 1-D bulk copy of the tile's contiguous interval
//                 (TILED) or a tensor-map copy of the interval's [nf rows x 32 segments] box (PLAIN).
// SHARED = true : forcings come from a small table of met-station series shared by all segments; each lane gathers
//                 its station's value of every row with an 8-byte cp.async (LDGSTS) into the same ring, one group per
//                 interval, and forcing() multiplies by the segment's MFACTOR at the point of use.  Lanes only ever
//                 touch their own ring column, so no barrier of any kind is involved.
template <bool SHARED, class LAY = TiledLay>
struct PipeT {
    unsigned char *ring;  // this warp's ring (generic pointer into shared memory)
    unsigned char *lane_ring;  // this lane's cell of row 0 of stage 0
    unsigned ring_s;      // its shared-space address
    unsigned bar_s;       // shared-space address of this warp's first mbarrier
    const LandLaunch *Lp;
    const unsigned char *src;  // SHARED: met + station of this lane; TILED: forcings of the current tile; PLAIN: column tile*32 of forc
    int stage_bytes;      // nf * 32 * element size
    int nf, nst;          // rows per interval; SHARED: stations
    int esz;              // forcing element size in bytes (8, or 4 for float32 forcings)
    int tile;             // PLAIN: current tile
    int src_tile;         // linked rows: the source batch's tile this tile reads
    int tbl;              // SHARED: 1 = one bulk copy of the interval's station table (L.met_table), 0 = per-lane gathers
    int my_station;       // SHARED table mode: this lane's station
    __device__ __forceinline__ bool table() const { return SHARED && (LAY::METTBL >= 0 ? LAY::METTBL != 0 : tbl != 0); }
    int t_end;            // end of the current task (no copies are issued past it)
    int st;               // stage of the current interval
    unsigned ph;          // its phase parity
    int lane;

    // smem layout: [warp][PIPE_STAGES*nf*32 ring | nhot*32 hot parameters (+ nf*32 MFACTORs when SHARED) | ncold*32 cold
    // state], then the mbarriers
    __device__ __forceinline__ void init(const LandLaunch &L, double *smem, int wib, int lane_, int nf_, int nhot, int ncold) {
        Lp = &L;
        nf = nf_;
        nst = L.n_stations;
        esz = (!SHARED && lay_f32in<LAY>(L)) ? 4 : 8;
        stage_bytes = nf * LAND_TILE * esz;
        const int warp_bytes = PIPE_STAGES * stage_bytes + (nhot + ncold) * LAND_TILE * 8;
        // byte offsets inside the block's shared array, made opaque (HSP_PIN32) so that they live in registers; the pointers are
        // formed from the shared array itself, which keeps every access an LDS / STS (a pinned POINTER turns them generic)
        unsigned ring_off = (unsigned)wib * (unsigned)warp_bytes;
        unsigned lane_off = ring_off + (unsigned)lane_ * (unsigned)esz;
        HSP_PIN32(ring_off);
        HSP_PIN32(lane_off);
        ring = reinterpret_cast<unsigned char *>(smem) + ring_off;
        lane_ring = reinterpret_cast<unsigned char *>(smem) + lane_off;
        ring_s = smem_u32(smem) + ring_off;
        bar_s = smem_u32(smem) + (unsigned)(LAND_WARPS * warp_bytes) + wib * PIPE_STAGES * 8;
        HSP_PIN32(ring_s);
        HSP_PIN32(bar_s);
        st = 0;
        ph = 0;
        lane = lane_;
        t_end = 0;
        tile = 0;
        src = nullptr;
        tbl = SHARED && lay_mettbl<LAY>(L);
        my_station = 0;
        if (!SHARED || tbl) {
            if (lane == 0) {
                for (int s = 0; s < PIPE_STAGES; ++s) mbar_init(bar_s + s * 8, 1);
                asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            }
            __syncwarp();
        }
    }
    __device__ __forceinline__ void issue(int t, int s) {  // SHARED gathers: every lane; else lane 0 only
        if (table()) {
            if (t < t_end) {
                const unsigned bytes = (unsigned)(nf * nst * 8);
                const unsigned bar = bar_s + s * 8;
                mbar_expect_tx(bar, bytes);
                bulk_g2s(ring_s + (unsigned)s * (unsigned)stage_bytes, src + (size_t)t * bytes, bytes, bar);
            }
        } else if (SHARED) {
            if (t < t_end) {
                double *dst = reinterpret_cast<double *>(ring + s * stage_bytes) + lane;
                const double *g = reinterpret_cast<const double *>(src) + (size_t)t * nf * nst;
                for (int r = 0; r < nf; ++r) cp_async8(dst + r * LAND_TILE, g + (size_t)r * nst);
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");  // empty groups keep the wait arithmetic uniform
        } else if (t < t_end) {
            const unsigned bytes = (unsigned)stage_bytes;
            const unsigned bar = bar_s + s * 8, dst = ring_s + (unsigned)s * bytes;
            mbar_expect_tx(bar, bytes);
            if (lay_linked<LAY>(*Lp)) {  // row by row, each from its own source buffer (src_tile: the source's tile)
                const unsigned rowb = (unsigned)(LAND_TILE * 8);
#pragma unroll 1
                for (int r = 0; r < nf; ++r) {
                    const RowSource &rs = Lp->rowsrc[r];
                    bulk_g2s(dst + r * rowb, rs.base + (long long)t * rs.t_stride + (long long)src_tile * rs.tile_stride, rowb, bar);
                }
            } else if (!lay_plain<LAY>(*Lp))
                bulk_g2s(dst, src + (size_t)t * bytes, bytes, bar);
            else if (Lp->tmap_ok)
                tensor_g2s(dst, Lp->tmap, tile * LAND_TILE, t, bar);
            else {  // row by row: nf bulk copies of one 32-segment row each
                const size_t ldb = (size_t)Lp->forc_ld * esz;
                const unsigned rowb = (unsigned)(LAND_TILE * esz);
#pragma unroll 1
                for (int r = 0; r < nf; ++r) bulk_g2s(dst + r * rowb, src + ((size_t)t * nf + r) * ldb, rowb, bar);
            }
        }
    }
    // a new task: intervals [t0, t1) of `tile`.  Every stage is free here (the previous task consumed all it issued), so the
    // first PIPE_STAGES intervals go out at once, starting at the current stage.
    __device__ __forceinline__ void begin_task(int tile_, int t0, int t1) {
        const LandLaunch &L = *Lp;
        tile = tile_;
        if (!SHARED && lay_linked<LAY>(L)) src_tile = tile % L.src_tiles;
        if (table()) {
            src = reinterpret_cast<const unsigned char *>(L.met);
            my_station = L.station[tile * LAND_TILE + lane];
        } else if (SHARED)
            src = reinterpret_cast<const unsigned char *>(L.met + L.station[tile * LAND_TILE + lane]);
        else if (lay_plain<LAY>(L))
            src = reinterpret_cast<const unsigned char *>(L.forc) + (size_t)tile * LAND_TILE * esz;
        else
            src = reinterpret_cast<const unsigned char *>(L.forc) + (size_t)tile * L.nsteps * stage_bytes;
        t_end = stage_bytes > 0 ? t1 : 0;
        if ((SHARED && !table()) || lane == 0) {
            int s = st;
            for (int k = 0; k < PIPE_STAGES; ++k) {
                issue(t0 + k, s);
                if (++s == PIPE_STAGES) s = 0;
            }
        }
    }
    // wait for the current interval's forcings; returns this lane's cell of row 0 of the stage
    __device__ __forceinline__ const unsigned char *acquire() {
        if (table()) {
            if (t_end > 0) mbar_wait(bar_s + st * 8, ph);
            return ring + st * stage_bytes + my_station * 8;  // this lane's station's cell of row 0 of the table
        }
        if (SHARED)
            asm volatile("cp.async.wait_group %0;\n" ::"n"(PIPE_STAGES - 1) : "memory");
        else if (t_end > 0)
            mbar_wait(bar_s + st * 8, ph);
        return lane_ring + st * stage_bytes;
    }
    __device__ __forceinline__ double *hot_area() const { return reinterpret_cast<double *>(ring + PIPE_STAGES * stage_bytes); }
    // all lanes are done with interval t: refill its stage with interval t + PIPE_STAGES
    __device__ __forceinline__ void release(int t) {
        if (SHARED && !table()) {
            issue(t + PIPE_STAGES, st);
        } else {
            __syncwarp();
            if (lane == 0) issue(t + PIPE_STAGES, st);
        }
        if (++st == PIPE_STAGES) {
            st = 0;
            ph ^= 1u;
        }
    }
    __device__ __forceinline__ void end_task() {
        if (SHARED && !table()) asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    }
};

// ---- task scheduler.  A launch covers nsteps intervals of every tile, cut into sub-chunks of L.sub intervals; a task
// is one (tile, sub-chunk).  Resident warps take tickets from one global counter, handed out sub-chunk-major, so
// the machine stays full until the last sub-chunk instead of idling through the tail of a second, partial wave of
// whole-chunk blocks (100k segments = 3,125 tiles against ~1,900 resident warps).  Sub-chunk c of a tile needs the
// state that c-1 left in the segment block: the warp that finished c-1 publishes it with a release store to
// progress[tile], the warp that drew c acquires it.  Tickets are drawn in order and every drawn ticket is held by a
// running warp, so the wait cannot deadlock.
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned *p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
// HSP_LOCKSTEP (kernels built per batch only, hspsquared_b200/jit.py `lockstep`): the warps of a block draw LAND_WARPS
// consecutive tickets at once (adjacent tiles of one sub-chunk, i.e. neighbours in the batch's climate order) and meet at a
// block barrier before every interval, so that the resident warps of an SM run the SAME stretch of the loop at about the same
// time and share its instruction-cache lines instead of sitting at 16 unrelated points of a 38 KB loop (profiles/README.md,
// the melt-season launch).  Every warp of a block passes exactly L.sub barriers per round, whether its own task is shorter
// (the last sub-chunk) or absent (tickets past the end): `idle`.  Dependencies (progress[tile]) point at tickets ntiles
// earlier, i.e. at earlier rounds, so the lowest unfinished round can always run: no deadlock.
struct Sched {
    const LandLaunch &L;
    int lane, c;
    // first thing a block does: let the next kernel of the stream be scheduled as soon as every block of this one has started.  It
    // matters only to a launch made with programmatic stream serialization (hsp2g_api.cu launch_jit, overlap mode): that launch's
    // blocks then take the places this grid's blocks free as they run out of tickets, instead of waiting for its last task.
    static __device__ __forceinline__ void release_dependents() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
#ifdef HSP_LOCKSTEP
    int wib;
    bool idle;
    __device__ __forceinline__ Sched(const LandLaunch &l, int lane_, int wib_) : L(l), lane(lane_), c(0), wib(wib_), idle(false) {
        release_dependents();
    }
#else
    __device__ __forceinline__ Sched(const LandLaunch &l, int lane_, int) : L(l), lane(lane_), c(0) { release_dependents(); }
#endif
    __device__ __forceinline__ bool next(int &tile, int &t0, int &t1) {
#ifdef HSP_LOCKSTEP
        __shared__ unsigned long long s_base;
        __syncthreads();  // the previous round's readers are done with s_base
        if (threadIdx.x == 0) s_base = atomicAdd(L.task_counter, (unsigned long long)LAND_WARPS) - L.task_base;
        __syncthreads();
        const unsigned long long base = s_base;
        if (base >= (unsigned long long)L.ntiles * (unsigned)L.nsub) return false;  // block-uniform
        const unsigned long long k = base + (unsigned)wib;
        idle = k >= (unsigned long long)L.ntiles * (unsigned)L.nsub;
        if (idle) {
            tile = 0;
            t0 = t1 = 0;
            return true;
        }
#else
        unsigned long long k = 0;
        if (lane == 0) k = atomicAdd(L.task_counter, 1ull) - L.task_base;
        k = __shfl_sync(0xffffffffu, k, 0);
        if (k >= (unsigned long long)L.ntiles * (unsigned)L.nsub) return false;
#endif
        c = (int)(k / (unsigned)L.ntiles);
        tile = (int)(k - (unsigned long long)c * (unsigned)L.ntiles);
        t0 = c * L.sub;
        t1 = t0 + L.sub < L.nsteps ? t0 + L.sub : L.nsteps;
        if (c > 0 || L.chained) {
            if (lane == 0)
                while ((int)(ld_acquire_u32(L.progress + tile) - (L.prog_base + (unsigned)c)) < 0) __nanosleep(256);
            __syncwarp();
        }
        return true;
    }
    __device__ __forceinline__ void done(int tile) {
#ifdef HSP_TASK_FENCE
        __threadfence();
#endif
        // every lane's state rows (weak L2 stores) are ordered before lane 0's release by the warp barrier, and the release
        // is cumulative at gpu scope: the warp that acquires progress[tile] and passes its own warp barrier sees them all.
        // (A __threadfence() by all 32 lanes here cost a second MEMBAR and an L1 invalidate per task.)
        __syncwarp();
        if (lane == 0) st_release_u32(L.progress + tile, L.prog_base + (unsigned)c + 1u);
    }
};
#define HSP_SYNCWARP() __syncwarp()
#ifdef HSP_LOCKSTEP
// a warp without a task keeps its block's barrier count; the time loop runs L.sub rounds for every warp of the block
#define HSP_TASK_IDLE(sched, L)                                   \
    if ((sched).idle) {                                           \
        for (int i_ = 0; i_ < (L).sub; ++i_) __syncthreads();     \
        continue;                                                 \
    }
#define HSP_T_LIM(L, t0, t1) ((t0) + (L).sub)
#define HSP_STEP_SYNC(t, t1) \
    __syncthreads();         \
    if ((t) >= (t1)) continue;
#if HSP_LOCKSTEP >= 2
#define HSP_SECTION_SYNC() __syncthreads()
#else
#define HSP_SECTION_SYNC()
#endif
#else
#define HSP_TASK_IDLE(sched, L)
#define HSP_T_LIM(L, t0, t1) (t1)
#define HSP_STEP_SYNC(t, t1)
#define HSP_SECTION_SYNC()
#endif

#else
// test double: same interface, synchronous copies (no TMA on the host); one "thread" at a time, which walks the
// sub-chunks of its own tile in order (no tickets)
template <bool SHARED, class LAY = TiledLay>
struct PipeT {
    unsigned char *ring;
    const LandLaunch *Lp;
    const unsigned char *src;
    int stage_bytes, lane, t_cur, nf, nst, esz, tile, src_tile, tbl, my_station;
    bool table() const { return SHARED && (LAY::METTBL >= 0 ? LAY::METTBL != 0 : tbl != 0); }
    void init(const LandLaunch &L, double *smem, int, int lane_, int nf_, int, int) {
        Lp = &L;
        ring = reinterpret_cast<unsigned char *>(smem);
        nf = nf_;
        nst = L.n_stations;
        tbl = SHARED && lay_mettbl<LAY>(L);
        my_station = 0;
        esz = (!SHARED && lay_f32in<LAY>(L)) ? 4 : 8;
        stage_bytes = nf * LAND_TILE * esz;
        lane = lane_;
    }
    void begin_task(int tile_, int t0, int) {
        const LandLaunch &L = *Lp;
        tile = tile_;
        if (!SHARED && lay_linked<LAY>(L)) src_tile = tile % L.src_tiles;
        if (table()) {
            src = reinterpret_cast<const unsigned char *>(L.met);
            my_station = L.station[tile * LAND_TILE + lane];
        } else if (SHARED)
            src = reinterpret_cast<const unsigned char *>(L.met + L.station[tile * LAND_TILE + lane]);
        else if (lay_plain<LAY>(L))
            src = reinterpret_cast<const unsigned char *>(L.forc) + (size_t)tile * LAND_TILE * esz;
        else
            src = reinterpret_cast<const unsigned char *>(L.forc) + (size_t)tile * L.nsteps * stage_bytes;
        t_cur = t0;
    }
    const unsigned char *acquire() {
        if (table()) {
            memcpy(ring, src + (size_t)t_cur * nf * nst * 8, (size_t)nf * nst * 8);
            return ring + my_station * 8;
        }
        if (SHARED) {
            const double *g = reinterpret_cast<const double *>(src);
            for (int r = 0; r < nf; ++r) reinterpret_cast<double *>(ring)[r * LAND_TILE + lane] = g[((size_t)t_cur * nf + r) * nst];
        } else if (lay_linked<LAY>(*Lp)) {
            for (int r = 0; r < nf; ++r) {
                const RowSource &rs = Lp->rowsrc[r];
                memcpy(ring + (size_t)r * LAND_TILE * 8, rs.base + (long long)t_cur * rs.t_stride + (long long)src_tile * rs.tile_stride,
                       (size_t)LAND_TILE * 8);
            }
        } else if (lay_plain<LAY>(*Lp)) {
            const size_t ldb = (size_t)Lp->forc_ld * esz;
            for (int r = 0; r < nf; ++r)
                memcpy(ring + (size_t)r * LAND_TILE * esz, src + ((size_t)t_cur * nf + r) * ldb, (size_t)LAND_TILE * esz);
        } else
            memcpy(ring, src + (size_t)t_cur * stage_bytes, (size_t)stage_bytes);
        return ring + lane * esz;
    }
    void release(int t) { t_cur = t + 1; }
    void end_task() {}
    double *hot_area() const { return reinterpret_cast<double *>(ring + PIPE_STAGES * stage_bytes); }
};
struct Sched {
    const LandLaunch &L;
    int c, wib;
    Sched(const LandLaunch &l, int, int wib_) : L(l), c(-1), wib(wib_) {}
    bool next(int &tile, int &t0, int &t1) {
        tile = (int)blockIdx.x * LAND_WARPS + wib;
        if (tile >= L.ntiles || ++c >= L.nsub) return false;
        t0 = c * L.sub;
        t1 = t0 + L.sub < L.nsteps ? t0 + L.sub : L.nsteps;
        return true;
    }
    void done(int) {}
};
#define HSP_SYNCWARP()
#define HSP_TASK_IDLE(sched, L)
#define HSP_T_LIM(L, t0, t1) (t1)
#define HSP_STEP_SYNC(t, t1)
#define HSP_SECTION_SYNC()
#endif

// ---- metric units (siminfo['units'] == 2; HSP2G_OPT_METRIC in L.opts).  The reference keeps every state in English units and
// converts (a) parameters and initial states on entry -- done by the host in the reference's expression order
// (hspsquared_b200/segstore.py) --, (b) some parameter SERIES on entry, elementwise after any monthly interpolation (CEPSC, UZSN,
// RETSC * 0.0394; NVSI * 2.205 / 2.471; ACCSDP * 1.10231 / 2.471; AWTF * 9 / 5 + 32: at the point of use) and the series one
// section hands the next (WYIELD, SURO, SURS, IFWO, AGWO: stored * 25.4 by the water section, read back * 0.0394 by SEDMNT /
// PWTGAS / SOLIDS / IWTGAS; the soil temperatures PSTEMP stores in deg C, land_kernels.cuh), and (c) the STORED series on exit
// (SNOW.py:569-585, PWATER.py:660-688, IWATER.py:276-283, SEDMNT.py:258-262, PWTGAS.py:338-350, SOLIDS.py:143-145,
// IWTGAS.py:179-182): `x * 25.4`, `x * 0.555 - 17.777` for the three snow temperatures, `x / 1.10231 * 2.471` (tons/ac ->
// tonnes/ha), `x / 3.96567 * 2.471` (heat), `x / 2.205 * 2.471` (mass).  PSTEMP / PWTGAS / IWTGAS temperatures: the sections
// themselves store deg C instead of converting to deg F.
#define OPT_METRIC 2u
__host__ __device__ constexpr int metric_out_code(int oid) {  // 0 as is, 1 x * 25.4, 2 x * 0.555 - 17.777, 3 / 4 / 5 see above
    return (oid == O_SNOW_PAKTMP || oid == O_SNOW_DEWTMP || oid == O_SNOW_SNOTMP) ? 2
           : (oid == O_SNOW_PACKF || oid == O_SNOW_PACKI || oid == O_SNOW_PACKW || oid == O_SNOW_PDEPTH || oid == O_SNOW_NEGHTS ||
              oid == O_SNOW_COVINX || oid == O_SNOW_SNOWE || oid == O_SNOW_SNOWF || oid == O_SNOW_WYIELD || oid == O_SNOW_XLNMLT ||
              oid == O_SNOW_MELT || oid == O_SNOW_PRAIN)
               ? 1
           : (oid >= O_PWATER_AGWET && oid <= O_PWATER_UZS && oid != O_PWATER_PETADJ && oid != O_PWATER_TGWS)                     ? 1
           : (oid >= O_IWATER_IMPEV && oid <= O_IWATER_SURS && oid != O_IWATER_PETADJ)                                             ? 1
           : (oid == O_SEDMNT_DETS || oid == O_SEDMNT_WSSD || oid == O_SEDMNT_SCRSD || oid == O_SEDMNT_SOSED ||
              oid == O_SOLIDS_SOSLD || oid == O_SOLIDS_SLDS)
               ? 3
           : (oid == O_PWTGAS_SOHT || oid == O_PWTGAS_IOHT || oid == O_PWTGAS_AOHT || oid == O_PWTGAS_POHT || oid == O_IWTGAS_SOHT) ? 4
           : (oid == O_PWTGAS_SODOXM || oid == O_PWTGAS_SOCO2M || oid == O_PWTGAS_IODOXM || oid == O_PWTGAS_IOCO2M ||
              oid == O_PWTGAS_AODOXM || oid == O_PWTGAS_AOCO2M || oid == O_PWTGAS_PODOXM || oid == O_PWTGAS_POCO2M ||
              oid == O_IWTGAS_SODOXM || oid == O_IWTGAS_SOCO2M)
               ? 5
               : 0;
}

// per-thread view of the batch

// Option words of the batch.  DynFlags reads each segment's 64-bit word from its segment block.  StaticFlags<W, MASK> fixes, at
// compile time, the bits every segment of the batch agrees on (MASK; their common values in W): the tests of those bits fold
// and the sections they switch off (degree-day snow, ARM routing, UZINF2, monthly-table paths ...) vanish from the kernel; only
// the bits that really differ between segments are read from the segment's word.  MASK = all ones is the replicated /
// calibration ensemble (one option word); an ensemble with a few options drawn per segment still agrees on most of the 62 bits.
// OPTS: the batch-wide HSP2G_OPT_* bits when they too are fixed at compile time (-1: read L.opts).
struct DynFlags {
    static constexpr bool KNOWN = false;
    static constexpr unsigned long long W = 0, MASK = 0;
    static constexpr int OPTS = -1;
};
template <unsigned long long W_, unsigned long long MASK_ = ~0ull, int OPTS_ = -1>
struct StaticFlags {
    static constexpr bool KNOWN = true;
    static constexpr unsigned long long W = W_, MASK = MASK_;
    static constexpr int OPTS = OPTS_;
};

template <class IO_, unsigned long long HOT_, bool SHARED_ = false, class FW_ = DynFlags, class LAY_ = TiledLay>
struct Seg {
    typedef IO_ IO;
    typedef FW_ FW;
    typedef LAY_ LAY;
    static constexpr bool SHARED = SHARED_;
    static constexpr unsigned long long HOT = HOT_;
    static constexpr int NHOT = ce_popc(HOT_);
    const LandLaunch &L;
    double *sb;            // this lane's column of the tile's segment block
    const double *hot;     // this lane's column of the staged hot parameters (shared memory)
    const double *mfacp;   // SHARED: this lane's column of the staged MFACTORs, one row per forcing row (shared memory)
    const double *monp;    // this lane's column of the tile's monthly tables
    unsigned char *outp;   // this lane's cell of row 0 of the current interval of out (float64 or float32 rows)
    const unsigned char *stage;  // this lane's cell of row 0 of the current forcing stage (shared memory; float64 or float32 rows)
    const CalStep *calp;   // calendar record of the current interval (warp-uniform)
    unsigned long long fl;
    unsigned calfl;        // HSP2G_CAL_* of the current interval
    unsigned calfl_next;   // ... of the next one, loaded an interval ahead (an L2 round trip off the critical path)
    long long out_step;    // bytes between consecutive intervals of out
    unsigned out_row;      // bytes between consecutive rows of out (PLAIN; TILED: 32 elements, a compile-time constant)
    int met_stride;        // SHARED: elements between consecutive rows of a forcing stage (n_stations in table mode, else 32)

    // view of `tile` for the task that starts at interval t0 of the launch
    __device__ __forceinline__ Seg(const LandLaunch &l, int tile, int lane, int t0, double *hot_area) : L(l) {
        sb = l.segblk + ((size_t)tile * SB_ROWS * LAND_TILE + lane);
        HSP_PIN(sb);
        monp = l.mon + ((size_t)tile * l.mon_rows * LAND_TILE + lane);
        const int osz = out_is_f32() ? 4 : 8;
        const int nsr = IO::STATIC ? IO::NS : l.ns;
        if (lay_plain<LAY>(l)) {
            out_row = (unsigned)(l.out_ld * osz);  // < 2^32: rows of fewer than 512 M segments
            HSP_PIN32(out_row);                    // an opaque 32-bit value: each store address is one IMAD.WIDE.U32
            out_step = (long long)out_row * nsr;
            outp = reinterpret_cast<unsigned char *>(l.out) + ((size_t)t0 * out_step + ((size_t)tile * LAND_TILE + lane) * osz);
        } else {
            out_row = LAND_TILE * osz;
            out_step = (long long)nsr * LAND_TILE * osz;
            outp = reinterpret_cast<unsigned char *>(l.out) + (((size_t)tile * l.nsteps + t0) * out_step + lane * osz);
        }
        HSP_PIN(outp);
        calp = l.cal + t0;
        double *h = hot_area + lane;
        int k = 0;
#pragma unroll
        for (int id = 0; id < 64 && id < HSP2G_NPAR; ++id)
            if ((HOT >> id) & 1ull) h[(k++) * LAND_TILE] = __ldg(sb + (SB_PAR + id) * LAND_TILE);
        hot = h;  // each lane reads back only what it wrote: no barrier needed
        mfacp = h + NHOT * LAND_TILE;
        if (SHARED) {
            const int nfr = IO::STATIC ? IO::NF : l.nf;
            const double *mg = l.mfac + ((size_t)tile * nfr * LAND_TILE + lane);
            for (int r = 0; r < nfr; ++r) h[(NHOT + r) * LAND_TILE] = __ldg(mg + r * LAND_TILE);
        }
        fl = *reinterpret_cast<const unsigned long long *>(sb + SB_FLAGS * LAND_TILE);
        met_stride = (SHARED && lay_mettbl<LAY>(l)) ? l.n_stations : LAND_TILE;
        stage = nullptr;
        calfl = 0;
        calfl_next = calp->flags;
    }
    __device__ __forceinline__ void begin_step(const unsigned char *stage_) {
        stage = stage_;
        calfl = calfl_next;
        calfl_next = calp[1].flags;  // the calendar holds steps+1 records (hsp2g_set_calendar), so t+1 exists
    }
    __device__ __forceinline__ void end_step() {
        outp += out_step;
        ++calp;
    }
    __device__ __forceinline__ bool out_is_f32() const { return IO::STATIC ? IO::F32 : (L.out_f32 != 0); }
    __device__ __forceinline__ void store_out(int r, double v) const {
        // TILED: rows at immediate offsets; PLAIN: rows out_row bytes apart, one widening multiply-add (IMAD.WIDE.U32) per store
        unsigned char *q = outp + ((LAY::STATIC && !LAY::PLAIN) ? (unsigned long long)(r * LAND_TILE * (out_is_f32() ? 4 : 8))
                                                                 : (unsigned long long)(unsigned)r * (unsigned long long)out_row);
        if (out_is_f32())
            HSP_STORE(reinterpret_cast<float *>(q), (float)v);
        else
            HSP_STORE(reinterpret_cast<double *>(q), v);
    }
    // forcing row r of the current interval (this lane's segment), widened when the series are float32
    __device__ __forceinline__ double stage_row(int r) const {
        if (!SHARED && lay_f32in<LAY>(L)) return (double)reinterpret_cast<const float *>(stage)[r * LAND_TILE];
        if (SHARED) return reinterpret_cast<const double *>(stage)[r * met_stride];  // station table: rows n_stations apart
        return reinterpret_cast<const double *>(stage)[r * LAND_TILE];
    }
    __device__ __forceinline__ bool flag(int bit) const { return ((((FW::MASK >> bit) & 1ull) ? FW::W : fl) >> bit) & 1ull; }
    __device__ __forceinline__ double par(int id) const {
        if (id < 64 && ((HOT >> (id & 63)) & 1ull)) return hot[hsp_popc(HOT & ((1ull << (id & 63)) - 1ull)) * LAND_TILE];
        return __ldg(sb + (SB_PAR + id) * LAND_TILE);
    }
    __device__ __forceinline__ bool has_forcing(int id) const {
        if (IO::STATIC) return (IO::FMASK >> id) & 1ull;
        return L.frow[id] >= 0;
    }
    __device__ __forceinline__ double forcing(int id) const {
        if (IO::STATIC) {
            if ((IO::FMASK >> id) & 1ull) {
                const int r = hsp_popc(IO::FMASK & ((1ull << id) - 1ull));
                return SHARED ? stage_row(r) * mfacp[r * LAND_TILE] : stage_row(r);
            }
            return L.ffill[id];
        }
        const int r = L.frow[id];
        if (r < 0) return L.ffill[id];
        return SHARED ? stage_row(r) * mfacp[r * LAND_TILE] : stage_row(r);
    }
    // monthly table value for the current day: numpy's slope*(x-x0)+y0 (calendar.py dayval), no FMA
    __device__ __forceinline__ double monthly(int mid) const {
        const double *tab = monp + (size_t)L.mslot[mid] * (12 * LAND_TILE);
        const double y0 = __ldg(tab + (int)calp->m0 * LAND_TILE);
        const double xm = calp->xm;
        if (xm == 0.0) return y0;
        const double y1 = __ldg(tab + (int)calp->m1 * LAND_TILE);
        const double slope = HDIV(y1 - y0, calp->dxm);
        return slope * xm + y0;
    }
    // N daily values at once: every table load is issued before the first division, so a new day costs one L2
    // round trip instead of N in series.  Same arithmetic as daily().
    template <int N>
    __device__ __forceinline__ void daily_batch(const int (&fb)[N], const int (&mid)[N], const int (&pid)[N],
                                                double (&v)[N]) const {
        double y0[N], y1[N];
        const int m0 = calp->m0, m1 = calp->m1;
        const double xm = calp->xm, dxm = calp->dxm;
#pragma unroll
        for (int k = 0; k < N; ++k) {
            if (flag(fb[k])) {
                const double *tab = monp + (size_t)L.mslot[mid[k]] * (12 * LAND_TILE);
                y0[k] = __ldg(tab + m0 * LAND_TILE);
                y1[k] = __ldg(tab + m1 * LAND_TILE);
            } else {
                y0[k] = par(pid[k]);
                y1[k] = y0[k];
            }
        }
#pragma unroll
        for (int k = 0; k < N; ++k) v[k] = (flag(fb[k]) && xm != 0.0) ? HDIV(y1[k] - y0[k], dxm) * xm + y0[k] : y0[k];
    }
    // initm(): the table where the segment's flag is on, else the constant parameter
    __device__ __forceinline__ double daily(int flagbit, int mid, int pid) const {
        return flag(flagbit) ? monthly(mid) : par(pid);
    }
    __device__ __forceinline__ bool saved(int oid) const {
        if (IO::STATIC) return io_saved<IO>(oid);
        return L.srow[oid] >= 0;
    }
    __device__ __forceinline__ unsigned opts() const { return FW::OPTS >= 0 ? (unsigned)FW::OPTS : L.opts; }
    __device__ __forceinline__ bool metric() const { return (opts() & OPT_METRIC) != 0; }
    __device__ __forceinline__ void save(int oid, double v) const {
        if (metric()) {  // the stored series only (the state stays in English units); a warp-uniform branch
            const int c = metric_out_code(oid);
            if (c == 1)
                v = v * 25.4;
            else if (c == 2)
                v = (v * 0.555) - 17.777;
            else if (c == 3)
                v = HDIV(v, 1.10231) * 2.471;
            else if (c == 4)
                v = HDIV(v, 3.96567) * 2.471;
            else if (c == 5)
                v = HDIV(v, 2.205) * 2.471;
        }
        if (IO::STATIC) {
            if (io_saved<IO>(oid)) store_out(io_row<IO>(oid), v);
        } else {
            const int r = L.srow[oid];
            if (r >= 0) store_out(r, v);
        }
    }
    // state rows are handed from SM to SM between the sub-chunks of a launch: L2-only accesses (L1s are not coherent)
    __device__ __forceinline__ double st(int sid) const { return __ldcg(sb + (SB_STATE + sid) * LAND_TILE); }
    __device__ __forceinline__ void st(int sid, double v) const { __stcg(sb + (SB_STATE + sid) * LAND_TILE, v); }
    __device__ __forceinline__ void count(int eid, int n) const {
        if (n) atomicAdd(reinterpret_cast<unsigned long long *>(sb) + (SB_ERR + eid) * LAND_TILE, (unsigned long long)n);
    }
};
