// hsp2g_api.cu -- C ABI of libhsp2g (include/hsp2g.h): batch handle, segment store in HBM, chunked runs.
//
// Streams: `comp` runs the kernels (state dependencies serialise there); `h2d` and `d2h` move host chunks through two
// device slots so that copy-in of chunk k+1, the kernel of chunk k and copy-out of chunk k-1 overlap (SURVEY 8b
// "Threading": one handle per GPU, internal streams).
#include "cuda_rt.h"
#ifdef HSP2G_HOST_EMU
#include <dlfcn.h>
#else
#include <cuda.h>
#endif

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/hsp2g.h"
#include "jit_info.h"
#include "land_common.cuh"
#include "launch.h"

static_assert(HSP2G_ATEMP == HSP2G_ACT_ATEMP && HSP2G_SNOW == HSP2G_ACT_SNOW && HSP2G_PWATER == HSP2G_ACT_PWATER &&
                  HSP2G_SEDMNT == HSP2G_ACT_SEDMNT && HSP2G_PSTEMP == HSP2G_ACT_PSTEMP &&
                  HSP2G_PWTGAS == HSP2G_ACT_PWTGAS && HSP2G_IWATER == HSP2G_ACT_IWATER &&
                  HSP2G_SOLIDS == HSP2G_ACT_SOLIDS && HSP2G_IWTGAS == HSP2G_ACT_IWTGAS &&
                  HSP2G_PQUAL == HSP2G_ACT_PQUAL && HSP2G_IQUAL == HSP2G_ACT_IQUAL,
              "public activity bits must match the kernel's");
static_assert(HSP2G_ABI_VERSION == 3, "bump _lib.py with the header");
static_assert(HSP2G_NFLAG <= 64, "flag word is 64 bits");
static_assert(HSP2G_OPT_METRIC == OPT_METRIC, "public option bits must match the kernel's");
static_assert(sizeof(CalStep) == 32, "CalStep layout");

static thread_local std::string g_err;

static int fail(const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}
#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) return fail("%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// A kernel built for one batch's exact configuration (hspsquared_b200/jit.py -> hsp2g_set_kernel_image)
struct JitKernel {
    bool on = false;
    unsigned long long info[JI_N] = {0};
    int resident = 0;  // blocks the device holds at once
    std::string name;
#ifdef HSP2G_HOST_EMU
    void *dl = nullptr;
    void (*fn)(const LandLaunch &) = nullptr;
#else
    CUmodule mod = nullptr;
    CUfunction fn = nullptr;
#endif
};

struct Slot {
    double *forc = nullptr, *out = nullptr;
    float *agg = nullptr;  // period statistics of the slot's chunk (hsp2g_set_output_periods)
    size_t cap_f = 0, cap_o = 0, cap_a = 0;
    cudaEvent_t h2d_done = nullptr, k_done = nullptr, d2h_done = nullptr;
};

struct hsp2g_batch {
    int device = 0, op = 0;
    int64_t n_seg = 0, n_pad = 0, ntiles = 0;
    unsigned act = 0, opts = 0;
    double delt = 60.0;
    double *segblk = nullptr;  // [tile][SB_ROWS][32]: option word, parameters, states, error counters
    double *mon = nullptr;     // [tile][slot*12 + month][32]
    std::vector<double> h_mon; // host copy [slot][12][n_seg] (tables arrive one at a time)
    CalStep *cal = nullptr;
    unsigned long long *d_counter = nullptr;  // [2] task tickets (land_common.cuh Sched); launches alternate between the two
    unsigned *d_progress = nullptr;           // [ntiles] sub-chunks completed
    unsigned long long task_base[2] = {0, 0};
    int ticket_slot = 0;
    unsigned prog_base = 0;
    // hsp2g_set_launch_overlap: a launch may start while the previous launch of this batch (same stream, nothing enqueued in
    // between) is still draining its last tasks
    int overlap = 0;
    cudaStream_t last_stream = nullptr;  // stream of the last launch
    bool last_recorded = false;          // last_launch has been recorded for it (overlap mode records lazily)
    bool last_full_grid = false;         // the last launch ran the batch's own kernel on a grid that fills the device
    int64_t chained_launches = 0;
    int sub_steps = 64;
    int out_f32 = 0;
    int plain = 0;     // series layout of run_device / run_host: 0 = TILED [tile][t][row][32], 1 = PLAIN [t][row][ld]
    int forc_f32 = 0;  // forcing series are float32
    JitKernel jit;
    int *d_periods = nullptr;  // device copy of `periods`
    cudaEvent_t group_fork = nullptr;   // hsp2g_run_device_group: the caller's stream at the time of the call
    cudaEvent_t last_launch = nullptr;  // recorded after every launch on the stream it used (state / error accessors wait for it)
    bool have_last = false;
    // shared-met mode
    double *met = nullptr;   // [n_met_steps][nf][n_stations]
    int *station = nullptr;  // [ntiles][32]
    double *mfac = nullptr;  // [ntiles][nf][32]
    int n_stations = 0;
    int64_t n_met_steps = 0;
    int64_t ncal = 0;
    int nmon_slots = 0;
    short mslot[HSP2G_NMON];
    short frow[HSP2G_NFORC];
    double ffill[HSP2G_NFORC];
    short srow[HSP2G_NOUT];
    int nf = -1, ns = -1;
    bool have_par = false, have_flags = false, have_state = false;
    std::vector<int> periods;  // hsp2g_set_output_periods: period of every interval of the run (empty = series out)
    int flags_uniform = 0;  // every segment carries flags_word (a replicated ensemble): kernels specialised on the word apply
    unsigned long long flags_word = 0;
    unsigned long long flags_and = 0, flags_or = 0;
    cudaStream_t comp = nullptr, h2d = nullptr, d2h = nullptr;
    Slot slot[2];
    int next_slot = 0;
    int64_t launches = 0;
    bool timing = false;
    double kernel_ms = 0.0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
    std::vector<cudaEvent_t> ev_pool;
    const char *kname = "";
};

namespace {
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

#define NAME_ITEM(n) #n ","
const char *k_params = HSP2G_PARAMS(NAME_ITEM);
const char *k_monthly = HSP2G_MONTHLY(NAME_ITEM);
const char *k_flags = HSP2G_FLAGS(NAME_ITEM);
const char *k_states = HSP2G_STATES(NAME_ITEM);
const char *k_forcings = HSP2G_FORCINGS(NAME_ITEM);
const char *k_errors = HSP2G_ERRORS(NAME_ITEM);
#define OA(n) "ATEMP_" #n ","
#define OS(n) "SNOW_" #n ","
#define OP(n) "PWATER_" #n ","
#define OD(n) "SEDMNT_" #n ","
#define OT(n) "PSTEMP_" #n ","
#define OG(n) "PWTGAS_" #n ","
#define OI(n) "IWATER_" #n ","
#define OL(n) "SOLIDS_" #n ","
#define OW(n) "IWTGAS_" #n ","
#define OQ(n) "PQUAL_" #n ","
#define OJ(n) "IQUAL_" #n ","
const char *k_outputs = HSP2G_OUT_ATEMP(OA) HSP2G_OUT_SNOW(OS) HSP2G_OUT_PWATER(OP) HSP2G_OUT_SEDMNT(OD)
    HSP2G_OUT_PSTEMP(OT) HSP2G_OUT_PWTGAS(OG) HSP2G_OUT_IWATER(OI) HSP2G_OUT_SOLIDS(OL) HSP2G_OUT_IWTGAS(OW)
    HSP2G_OUT_PQUAL(OQ) HSP2G_OUT_IQUAL(OJ);

// host [rows][ld] (segment-minor) <-> device tile-major rows [tile][row0 + r][32] of a block with `blk_rows` rows per
// tile.  Padded lanes of the last tile receive copies of the last real segment (land_common.cuh).
// A synchronous cudaMemcpy FROM PAGEABLE host memory returns once the source is staged; the DMA into device memory may still be
// in flight (CUDA "API synchronization behavior").  The batch's streams are non-blocking -- they do not order themselves with the
// default stream -- so every such upload is followed by a wait on the default stream before anything can be launched on them.
#define H2D_DONE() CU(cudaStreamSynchronize(0))

int upload_rows(double *blk, int blk_rows, int row0, const void *src, int64_t ld, int64_t n_seg, int64_t ntiles,
                int rows) {
    std::vector<uint64_t> pk((size_t)ntiles * rows * LAND_TILE);
    const uint64_t *in = (const uint64_t *)src;
    for (int64_t t = 0; t < ntiles; ++t)
        for (int r = 0; r < rows; ++r)
            for (int l = 0; l < LAND_TILE; ++l) {
                int64_t seg = t * LAND_TILE + l;
                if (seg >= n_seg) seg = n_seg - 1;
                pk[((size_t)t * rows + r) * LAND_TILE + l] = in[(size_t)r * ld + seg];
            }
    const size_t w = (size_t)rows * LAND_TILE * 8;
    CU(cudaMemcpy2D(blk + (size_t)row0 * LAND_TILE, (size_t)blk_rows * LAND_TILE * 8, pk.data(), w, w, (size_t)ntiles,
                    cudaMemcpyHostToDevice));
    H2D_DONE();
    return 0;
}
int download_rows(void *dst, int64_t ld, const double *blk, int blk_rows, int row0, int64_t n_seg, int64_t ntiles,
                  int rows) {
    std::vector<uint64_t> pk((size_t)ntiles * rows * LAND_TILE);
    const size_t w = (size_t)rows * LAND_TILE * 8;
    CU(cudaMemcpy2D(pk.data(), w, blk + (size_t)row0 * LAND_TILE, (size_t)blk_rows * LAND_TILE * 8, w, (size_t)ntiles,
                    cudaMemcpyDeviceToHost));
    uint64_t *out = (uint64_t *)dst;
    for (int64_t t = 0; t < ntiles; ++t)
        for (int r = 0; r < rows; ++r)
            for (int l = 0; l < LAND_TILE; ++l) {
                const int64_t seg = t * LAND_TILE + l;
                if (seg < n_seg) out[(size_t)r * ld + seg] = pk[((size_t)t * rows + r) * LAND_TILE + l];
            }
    return 0;
}

int check_ready(const hsp2g_batch *b, int64_t t0, int32_t nsteps) {
    if (!b) return fail("null batch");
    if (!b->have_par) return fail("hsp2g_set_params has not been called");
    if (!b->have_flags) return fail("hsp2g_set_flags has not been called");
    if (!b->have_state) return fail("hsp2g_set_state has not been called");
    if (!b->cal) return fail("hsp2g_set_calendar has not been called");
    if (b->nf < 0) return fail("hsp2g_set_forcing_layout has not been called");
    if (b->ns < 0) return fail("hsp2g_set_save has not been called");
    if (nsteps <= 0) return fail("nsteps must be positive");
    if (t0 < 0 || t0 + nsteps > b->ncal) return fail("intervals [%lld, %lld) outside the calendar (%lld records)",
                                                   (long long)t0, (long long)(t0 + nsteps), (long long)b->ncal);
    return 0;
}

// the batch's kernels may run on a caller's stream (hsp2g_run_device): state and error rows are read / written only after the
// last launch has finished, whatever stream it used
int ensure_last_recorded(hsp2g_batch *b) {
    if (b->have_last && !b->last_recorded) {
        CU(cudaEventRecord(b->last_launch, b->last_stream));
        b->last_recorded = true;
    }
    return 0;
}
int wait_last_launch(hsp2g_batch *b) {
    CU(cudaStreamSynchronize(b->comp));
    if (ensure_last_recorded(b)) return 1;
    if (b->have_last) CU(cudaEventSynchronize(b->last_launch));
    return 0;
}

void drop_kernel_image(hsp2g_batch *b);

cudaEvent_t get_event(hsp2g_batch *b) {
    if (!b->ev_pool.empty()) {
        cudaEvent_t e = b->ev_pool.back();
        b->ev_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

int drain_timing(hsp2g_batch *b) {
    for (auto &p : b->pending) {
        CU(cudaEventSynchronize(p.second));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, p.first, p.second));
        b->kernel_ms += ms;
        b->ev_pool.push_back(p.first);
        b->ev_pool.push_back(p.second);
    }
    b->pending.clear();
    return 0;
}

#ifndef HSP2G_HOST_EMU
// The few driver-API entry points the library needs (module loading for kernels built at run time, tensor maps for the PLAIN
// layout), resolved through the runtime so that libhsp2g.so does not link libcuda and still loads on a box without a driver.
struct Driver {
    bool ok = false;
    std::string why;
    CUresult (*ModuleLoadData)(CUmodule *, const void *) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction *, CUmodule, const char *) = nullptr;
    CUresult (*ModuleGetGlobal)(CUdeviceptr *, size_t *, CUmodule, const char *) = nullptr;
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int *, CUfunction, int, size_t) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **,
                             void **) = nullptr;
    CUresult (*LaunchKernelEx)(const CUlaunchConfig *, CUfunction, void **, void **) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char **) = nullptr;
    CUresult (*TensorMapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                     const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill) = nullptr;
};
const Driver &driver() {
    static Driver d;
    static std::once_flag once;
    std::call_once(once, [] {
        struct {
            const char *name;
            void **slot;
        } want[] = {{"cuModuleLoadData", (void **)&d.ModuleLoadData},
                    {"cuModuleUnload", (void **)&d.ModuleUnload},
                    {"cuModuleGetFunction", (void **)&d.ModuleGetFunction},
                    {"cuModuleGetGlobal", (void **)&d.ModuleGetGlobal},
                    {"cuFuncSetAttribute", (void **)&d.FuncSetAttribute},
                    {"cuOccupancyMaxActiveBlocksPerMultiprocessor", (void **)&d.OccupancyMaxActiveBlocksPerMultiprocessor},
                    {"cuLaunchKernel", (void **)&d.LaunchKernel},
                    {"cuLaunchKernelEx", (void **)&d.LaunchKernelEx},
                    {"cuGetErrorString", (void **)&d.GetErrorString},
                    {"cuTensorMapEncodeTiled", (void **)&d.TensorMapEncodeTiled}};
        for (auto &w : want) {
            cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
            cudaError_t e = cudaGetDriverEntryPoint(w.name, w.slot, cudaEnableDefault, &q);
            if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !*w.slot) {
                d.why = std::string("driver entry point ") + w.name + " unavailable";
                return;
            }
        }
        d.ok = true;
    });
    return d;
}
const char *cu_str(CUresult r) {
    const char *m = nullptr;
    if (driver().GetErrorString) driver().GetErrorString(r, &m);
    return m ? m : "unknown driver error";
}
// CUtensorMap of a PLAIN forcing buffer [nsteps][nf][ld] (elements of esz bytes) as the 3-D tensor (segment, row, interval)
// with a box of 32 segments x nf rows x 1 interval: what one warp needs for one interval, fetched by one TMA instruction.
bool make_tmap(unsigned long long *out16, const void *base, int esz, long long ld, int nf, long long nsteps) {
    const Driver &d = driver();
    if (!d.ok || nf <= 0 || nf > 256 || ((uintptr_t)base & 15) || ((ld * esz) & 15)) return false;
    alignas(64) CUtensorMap m;
    const cuuint64_t dim[3] = {(cuuint64_t)ld, (cuuint64_t)nf, (cuuint64_t)nsteps};
    const cuuint64_t stride[2] = {(cuuint64_t)ld * esz, (cuuint64_t)ld * esz * nf};
    const cuuint32_t box[3] = {LAND_TILE, (cuuint32_t)nf, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    CUresult r = d.TensorMapEncodeTiled(&m, esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3,
                                        const_cast<void *>(base), dim, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap is 128 bytes");
    memcpy(out16, &m, sizeof m);
    return true;
}
#endif

void drop_kernel_image(hsp2g_batch *b) {
    if (!b) return;
#ifdef HSP2G_HOST_EMU
    if (b->jit.dl) dlclose(b->jit.dl);
    b->jit.dl = nullptr;
#else
    if (b->jit.mod && driver().ok) driver().ModuleUnload(b->jit.mod);
    b->jit.mod = nullptr;
#endif
    b->jit.fn = nullptr;
    b->jit.on = false;
}

// Does the kernel built for this batch (hsp2g_set_kernel_image) still describe the launch?  Anything that was fixed at
// compile time must equal the batch's current configuration; otherwise the library's general kernels run.
bool jit_matches(const hsp2g_batch *b, const LandLaunch &L) {
    const unsigned long long *I = b->jit.info;
    const unsigned long long bits = I[JI_BITS];
    const bool qual = (b->act & (HSP2G_ACT_PQUAL | HSP2G_ACT_IQUAL)) != 0;
    if (I[JI_OP] != (unsigned long long)b->op || ((bits & JB_QUAL) != 0) != qual) return false;
    if (!qual && I[JI_ACT] != ~0ull && I[JI_ACT] != b->act) return false;
    if (I[JI_HOURLY] && !(b->delt >= 60.0)) return false;
    if (((bits & JB_SHARED) != 0) != (L.met != nullptr)) return false;
    if (bits & JB_WORD) {
        // every bit fixed at compile time must be one all segments agree on, with the value the kernel was built for
        const unsigned long long mask = I[JI_MASK];
        if (((L.flags_and ^ L.flags_or) & mask) || ((L.flags_and ^ I[JI_WORD]) & mask)) return false;
    }
    if ((bits & JB_OPTS) && I[JI_OPTS] != (unsigned long long)b->opts) return false;  // HSP2G_OPT_* folded into the kernel
    if (bits & JB_LAYSTATIC) {
        if (((bits & JB_PLAIN) != 0) != (L.plain != 0)) return false;
        if (!(bits & JB_SHARED) && ((bits & JB_F32IN) != 0) != (L.forc_f32 != 0)) return false;
        if ((bits & JB_SHARED) && ((bits & JB_METTBL) != 0) != (L.met_table != 0)) return false;
        if (((bits & JB_LINKED) != 0) != (L.linked != 0)) return false;
    }
    if (bits & JB_IOSTATIC) {
        if (((bits & JB_F32OUT) != 0) != (L.out_f32 != 0)) return false;
        int r = 0;
        for (int id = 0; id < HSP2G_NFORC; ++id) {
            const bool on = (I[JI_FMASK] >> id) & 1ull;
            if (L.frow[id] != (on ? r : -1)) return false;
            r += on;
        }
        if (r != L.nf) return false;
        r = 0;
        for (int id = 0; id < HSP2G_NOUT; ++id) {
            const bool on = (I[JI_S0 + id / 64] >> (id & 63)) & 1ull;
            if (L.srow[id] != (on ? r : -1)) return false;
            r += on;
        }
        if (r != L.ns) return false;
    }
    return true;
}

// may_chain: the previous launch of this batch went to the same stream on a full grid, and the caller allows overlap.
// *full_grid: this launch fills the device (grid = resident blocks), so a chained successor can only become resident as this
// launch's blocks exit: at most two launches of the batch are ever in flight, one per ticket counter.
int launch_jit(hsp2g_batch *b, LandLaunch &L, cudaStream_t st, int *grid, bool may_chain, bool *full_grid, bool *chained) {
    const unsigned block = (unsigned)b->jit.info[JI_BLOCK];  // the image's own block size (jit.py `block`)
    // task length: the batch's (hsp2g_set_schedule, default 64).  Measured with overlapping launches on the headline alone, longer
    // tasks win (438-interval launches of 100,000 segments: 64 -> 2.135 ms, 110 -> 2.105, 146 -> 2.090, 219 -> 2.083,
    // profiles/r3k_overlap_task_length.jsonl) -- but a full bench.py run with "half a launch" as the default for every configuration
    // did not finish (cause not found before the round's GPU budget ran out), so the default stays at the validated 64.
    hsp2g_plan_tasks(L, b->sub_steps, b->jit.resident, (int)block / LAND_TILE, grid);
    *full_grid = *grid == b->jit.resident;
    // a launch of whole-chunk tasks (one task per tile) is never chained: measured to stall on the B200 (bench.py --sub 438 did not
    // finish in 200 s); it waits for its predecessor to end like any other kernel of the stream
    *chained = may_chain && *full_grid && L.nsub >= 2;
    L.chained = b->overlap ? 1 : 0;  // sub-chunk 0 of a tile also waits for the tile's previous launch (progress[tile])
#ifdef HSP2G_HOST_EMU
    (void)st;
    for (unsigned blk = 0; blk < (unsigned)*grid; ++blk)
        for (unsigned t = 0; t < block; ++t) {
            blockIdx.x = blk;
            threadIdx.x = t;
            b->jit.fn(L);
        }
#else
    void *params[1] = {&L};
    CUresult r;
    if (*chained) {
        // programmatic dependent launch: the blocks of this grid may become resident as soon as every block of the previous
        // kernel in the stream has started (ours call griddepcontrol.launch_dependents first thing), i.e. while that kernel's last
        // tasks are still running.  Ordering between the two launches is per tile, through progress[tile] (land_common.cuh Sched).
        CUlaunchAttribute attr[1];
        memset(attr, 0, sizeof attr);
        attr[0].id = CU_LAUNCH_ATTRIBUTE_PROGRAMMATIC_STREAM_SERIALIZATION;
        attr[0].value.programmaticStreamSerializationAllowed = 1;
        CUlaunchConfig cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDimX = (unsigned)*grid;
        cfg.gridDimY = cfg.gridDimZ = 1;
        cfg.blockDimX = block;
        cfg.blockDimY = cfg.blockDimZ = 1;
        cfg.sharedMemBytes = (unsigned)b->jit.info[JI_SMEM];
        cfg.hStream = (CUstream)st;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        r = driver().LaunchKernelEx(&cfg, b->jit.fn, params, nullptr);
    } else
        r = driver().LaunchKernel(b->jit.fn, (unsigned)*grid, 1, 1, block, 1, 1, (unsigned)b->jit.info[JI_SMEM], (CUstream)st, params,
                                  nullptr);
    if (r != CUDA_SUCCESS) return fail("launch of the batch's own kernel failed: %s", cu_str(r));
#endif
    b->kname = b->jit.name.c_str();
    return 0;
}

// One launch over intervals [t0, t0+nsteps).  Series layout: b->plain ? PLAIN with rows forc_ld / out_ld elements apart : TILED.
int launch(hsp2g_batch *b, int64_t t0, int32_t nsteps, const void *d_forc, long long forc_ld, void *d_out, long long out_ld,
           cudaStream_t st, const hsp2g_row_source *rows = nullptr, long long src_tiles = 0) {
    LandLaunch L;
    memset(&L, 0, sizeof L);
    if (rows) {  // hsp2g_run_device_linked: every forcing row is read in place from another batch's device buffer
        L.linked = 1;
        L.src_tiles = (int)src_tiles;
        for (int r = 0; r < b->nf; ++r) {
            L.rowsrc[r].base = (const unsigned char *)rows[r].base;
            L.rowsrc[r].t_stride = rows[r].interval_stride;
            L.rowsrc[r].tile_stride = rows[r].tile_stride;
        }
    }
    L.segblk = b->segblk;
    L.mon = b->mon;
    L.forc = (const double *)d_forc;
    L.out = (double *)d_out;
    if (b->met) {
        if (t0 + nsteps > b->n_met_steps)
            return fail("intervals [%lld, %lld) outside the shared met series (%lld intervals)", (long long)t0,
                        (long long)(t0 + nsteps), (long long)b->n_met_steps);
        L.met = b->met + (size_t)t0 * b->nf * b->n_stations;
        L.station = b->station;
        L.mfac = b->mfac;
        L.n_stations = b->n_stations;
        // opt-in (HSP2G_MET_TABLE=1): measured 4.5 % SLOWER than the per-lane gathers on B200 (profiles/README.md, round 2)
        const char *tbl = getenv("HSP2G_MET_TABLE");
        L.met_table = tbl && tbl[0] == '1' && met_table_fits(b->n_stations, b->nf);
    }
    L.cal = b->cal + t0;
    L.ntiles = (int)b->ntiles;
    L.mon_rows = 12 * b->nmon_slots;
    L.nsteps = nsteps;
    L.first_chunk = t0 == 0;
    L.nf = b->nf;
    L.ns = b->ns;
    L.out_f32 = b->out_f32;
    L.plain = b->plain;
    L.forc_f32 = b->met ? 0 : b->forc_f32;
    L.forc_ld = forc_ld;
    L.out_ld = out_ld;
    L.act = b->act;
    L.opts = b->opts;
    L.flags_uniform = b->flags_uniform;
    L.flags_word = b->flags_word;
    L.flags_and = b->flags_and;
    L.flags_or = b->flags_or;
    L.delt = b->delt;
    L.delt60 = b->delt / 60.0;
    memcpy(L.frow, b->frow, sizeof L.frow);
    memcpy(L.ffill, b->ffill, sizeof L.ffill);
    memcpy(L.srow, b->srow, sizeof L.srow);
    memcpy(L.mslot, b->mslot, sizeof L.mslot);
#ifndef HSP2G_HOST_EMU
    if (b->plain && !b->met && !rows && b->nf > 0 && !getenv("HSP2G_NO_TMAP"))
        L.tmap_ok = make_tmap(L.tmap, d_forc, L.forc_f32 ? 4 : 8, forc_ld, b->nf, nsteps) ? 1 : 0;
#endif
    const bool hourly = b->delt >= 60.0;
    // launches of one batch share its state rows: each waits for the previous one, whatever its stream (stream order does it when
    // the stream is the same; in overlap mode the event is recorded only when somebody needs it)
    if (b->have_last && !(b->overlap && b->last_stream == st)) {
        if (ensure_last_recorded(b)) return 1;
        CU(cudaStreamWaitEvent(st, b->last_launch, 0));
    }
    const bool may_chain = b->overlap && b->have_last && b->last_stream == st && b->last_full_grid && !b->last_recorded && !b->timing;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (b->timing) {
        e0 = get_event(b);
        e1 = get_event(b);
        CU(cudaEventRecord(e0, st));
    }
    const int slot = b->ticket_slot;
    L.task_counter = b->d_counter + slot;
    L.task_base = b->task_base[slot];
    L.progress = b->d_progress;
    L.prog_base = b->prog_base;
    int grid = 0;
    const bool qual = (b->act & (HSP2G_ACT_PQUAL | HSP2G_ACT_IQUAL)) != 0;
    int rc = 0;
    long long warps_per_block = LAND_WARPS;
    bool lockstep = false;
    bool full_grid = false, chained = false;
    if (b->jit.on && jit_matches(b, L)) {
        rc = launch_jit(b, L, st, &grid, may_chain, &full_grid, &chained);
        warps_per_block = (long long)b->jit.info[JI_BLOCK] / LAND_TILE;
        lockstep = (b->jit.info[JI_BITS] & JB_LOCKSTEP) != 0;
    } else {
        cudaError_t e = qual                    ? hsp2g_launch_qual(L, b->op == HSP2G_IMPLND, b->sub_steps, st, &b->kname, &grid)
                        : b->op == HSP2G_PERLND ? hsp2g_launch_perlnd(L, hourly, b->sub_steps, st, &b->kname, &grid)
                                                : hsp2g_launch_implnd(L, hourly, b->sub_steps, st, &b->kname, &grid);
        if (e != cudaSuccess) rc = fail("kernel launch failed: %s", cudaGetErrorString(e));
    }
    if (rc) {
        if (e0) b->ev_pool.push_back(e0);
        if (e1) b->ev_pool.push_back(e1);
        return rc;
    }
    // every task draws one ticket and every warp one more (the draw that tells it to stop)
    // (lock-step kernels draw a block's tickets together: whole rounds of warps_per_block, and one last round per block)
    const unsigned long long tasks = (unsigned long long)L.ntiles * L.nsub;
    b->task_base[slot] += (lockstep ? (tasks + warps_per_block - 1) / warps_per_block * warps_per_block : tasks) +
                          (unsigned long long)grid * warps_per_block;
    b->ticket_slot ^= 1;
    b->prog_base += (unsigned)L.nsub;
    if (b->timing) {
        CU(cudaEventRecord(e1, st));
        b->pending.emplace_back(e0, e1);
    }
    // the state / error accessors and the next launch of this batch order themselves after this one, whatever stream it is on
    b->last_stream = st;
    b->last_full_grid = full_grid;
    b->chained_launches += chained ? 1 : 0;
    b->last_recorded = false;
    b->have_last = true;
    if (!b->overlap && ensure_last_recorded(b)) return 1;
    b->launches += 1;
    return 0;
}
}  // namespace

#ifndef HSP2G_HOST_EMU
__global__ void math_probe_kernel(int fn, long long n, const double *x, const double *y, double *out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = fn == 0 ? hsp_pow(x[i], y[i]) : fn == 1 ? hsp_exp(x[i]) : fn == 2 ? hsp_exp2(x[i]) : hsp_log(x[i]);
}
#endif

// ---- row gather between tile-major series buffers (hsp2g_gather_*): destination entity j = (dtile, lane) takes its rows from
// source segment seg_of[j].  One warp per destination tile walks the intervals; per interval and row a lane moves 8 bytes
// (neighbouring lanes read neighbouring segments' cells of the same 256-byte source row, so the reads stay sector-efficient).
#define GATHER_MAXROWS 16
struct GatherArgs {
    const double *src;
    double *dst;
    const long long *seg_of;
    long long n_dst_pad;
    int nsteps, src_rows, dst_rows, k;
    short src_row[GATHER_MAXROWS], dst_row[GATHER_MAXROWS];
};
#ifndef HSP2G_HOST_EMU
__global__ void __launch_bounds__(128) gather_kernel(const __grid_constant__ GatherArgs A) {
    const long long j = (long long)blockIdx.x * 128 + threadIdx.x;
    if (j >= A.n_dst_pad) return;
    const long long seg = A.seg_of[j];
    const double *s = A.src + ((size_t)(seg / LAND_TILE) * A.nsteps * A.src_rows) * LAND_TILE + (seg % LAND_TILE);
    double *d = A.dst + ((size_t)(j / LAND_TILE) * A.nsteps * A.dst_rows) * LAND_TILE + (j % LAND_TILE);
    for (int t = 0; t < A.nsteps; ++t) {
        double v[GATHER_MAXROWS];
#pragma unroll
        for (int r = 0; r < GATHER_MAXROWS; ++r)
            if (r < A.k) v[r] = __ldcs(s + ((size_t)t * A.src_rows + A.src_row[r]) * LAND_TILE);
#pragma unroll
        for (int r = 0; r < GATHER_MAXROWS; ++r)
            if (r < A.k) d[((size_t)t * A.dst_rows + A.dst_row[r]) * LAND_TILE] = v[r];
    }
}
#endif

// ---- period aggregation of saved float32 series (hsp2g_aggregate_f32): what IOManager.write_ts does with pandas for
// outstep 3 / 4 / 5 (hsp2io/io.py:74-94: resample(...).last() / .sum() / .mean() of the float32 frame).  pandas' group_sum /
// group_mean run a compensated (Kahan) sum IN float32 over the non-NaN values in arrival order, group_last takes the last non-NaN
// value; the same recurrences here, one thread per (tile, row, lane) walking the chunk's intervals.  bin[t] (which period an
// interval belongs to) comes from the host, which asks pandas itself for the binning.
struct AggArgs {
    const float *src;
    float *dst;
    const int *bin;       // period of every interval, bin[0] belongs to the first interval of the chunk
    int bin_base;         // subtracted from bin[t]
    long long n_threads;  // TILED: ntiles * ns * 32; PLAIN: ns * ld
    long long ld;         // PLAIN: elements per row of src and dst
    int nsteps, ns, nbins, plain;
};
__host__ __device__ inline void agg_thread(const AggArgs &A, long long j) {
    // this thread's series: element t at s[t * sstep]; its three statistics of period p at d[p * dstep + k * kstep]
    const float *s;
    float *d;
    size_t sstep, dstep, kstep;
    if (A.plain) {
        const long long seg = j % A.ld, r = j / A.ld;
        s = A.src + (size_t)r * A.ld + seg;
        sstep = (size_t)A.ns * A.ld;
        d = A.dst + ((size_t)r * 3) * A.ld + seg;
        dstep = (size_t)A.ns * 3 * A.ld;
        kstep = (size_t)A.ld;
    } else {
        const int lane = (int)(j % LAND_TILE);
        const long long tr = j / LAND_TILE;
        const int r = (int)(tr % A.ns);
        const long long tile = tr / A.ns;
        s = A.src + (((size_t)tile * A.nsteps) * A.ns + r) * LAND_TILE + lane;
        sstep = (size_t)A.ns * LAND_TILE;
        d = A.dst + ((((size_t)tile * A.nbins) * A.ns + r) * 3) * LAND_TILE + lane;
        dstep = (size_t)A.ns * 3 * LAND_TILE;
        kstep = LAND_TILE;
    }
    const float nanf_ = __builtin_nanf("");
#define AGG_FLUSH(p, last_, sum_, cnt_)                                                            \
    do {                                                                                           \
        float *q_ = d + (size_t)(p) * dstep;                                                       \
        q_[0] = (last_);                                                                           \
        q_[kstep] = (sum_);                             /* min_count = 0: an all-NaN period sums to 0 */ \
        q_[2 * kstep] = (cnt_) ? (sum_) / (float)(cnt_) : nanf_; /* mean = compensated sum / count, in float32 */ \
    } while (0)
    float sum = 0.f, comp = 0.f, last = nanf_;
    int cnt = 0, cur = A.bin[0] - A.bin_base;
    for (int b = 0; b < cur; ++b) AGG_FLUSH(b, nanf_, 0.f, 0);  // leading empty periods
    for (int t = 0; t < A.nsteps; ++t) {
        const int b = A.bin[t] - A.bin_base;
        if (b != cur) {
            AGG_FLUSH(cur, last, sum, cnt);
            for (int e = cur + 1; e < b; ++e) AGG_FLUSH(e, nanf_, 0.f, 0);
            sum = 0.f, comp = 0.f, last = nanf_, cnt = 0, cur = b;
        }
        const float v = s[(size_t)t * sstep];
        if (v == v) {  // skipna
            ++cnt;
            last = v;
            const float y = v - comp;
            const float tt = sum + y;
            comp = (tt - sum) - y;
            if (comp != comp) comp = 0.f;  // pandas resets the compensation after an inf
            sum = tt;
        }
    }
    AGG_FLUSH(cur, last, sum, cnt);
    for (int e = cur + 1; e < A.nbins; ++e) AGG_FLUSH(e, nanf_, 0.f, 0);
#undef AGG_FLUSH
}
#ifndef HSP2G_HOST_EMU
__global__ void __launch_bounds__(128) aggregate_kernel(const __grid_constant__ AggArgs A) {
    const long long j = (long long)blockIdx.x * 128 + threadIdx.x;
    if (j < A.n_threads) agg_thread(A, j);
}
#endif
// enqueue the aggregation of one chunk; `bin` is a DEVICE array in the product build (a host array in the test double)
static int run_aggregate(const AggArgs &A, cudaStream_t st) {
#ifdef HSP2G_HOST_EMU
    (void)st;
    for (long long j = 0; j < A.n_threads; ++j) agg_thread(A, j);
    return 0;
#else
    aggregate_kernel<<<(unsigned)((A.n_threads + 127) / 128), 128, 0, st>>>(A);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("aggregate launch failed: %s", cudaGetErrorString(e));
    return 0;
#endif
}

// Kernels opt in to their dynamic shared memory once (the per-warp TMA rings can exceed the 48 KB default when many
// forcing rows are present); the resident-block count sizes the persistent grid.
cudaError_t hsp2g_prepare_kernel(const void *kernel, size_t smem_bytes, int block, int *resident) {
    static std::mutex mu;
    static std::map<const void *, size_t> done;
    std::lock_guard<std::mutex> lk(mu);
    auto it = done.find(kernel);
    if (it == done.end() || it->second < smem_bytes) {
        if (smem_bytes > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
            if (e != cudaSuccess) return e;
        }
        done[kernel] = smem_bytes;
    }
    int per_sm = 0, dev = 0, sms = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem_bytes);
    if (e != cudaSuccess) return e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    *resident = per_sm * sms;
    return *resident > 0 ? cudaSuccess : cudaErrorLaunchOutOfResources;
}

void hsp2g_plan_tasks(LandLaunch &L, int sub_pref, int resident_blocks, int warps_per_block, int *grid_out) {
    const long long resident = (long long)resident_blocks * warps_per_block;  // warps the device holds at once
    if (L.ntiles <= resident || sub_pref <= 0 || sub_pref >= L.nsteps) {
        L.sub = L.nsteps;  // every tile is resident at once (or sub-chunking is off): one task per tile
        L.nsub = 1;
    } else {
        L.sub = sub_pref;
        L.nsub = (L.nsteps + sub_pref - 1) / sub_pref;
    }
    const long long tasks = (long long)L.ntiles * L.nsub;
#ifdef HSP2G_HOST_EMU
    (void)tasks;
    *grid_out = (L.ntiles + warps_per_block - 1) / warps_per_block;  // the test double walks a tile's sub-chunks in one "thread"
#else
    const long long warps = tasks < resident ? tasks : resident;
    *grid_out = (int)((warps + warps_per_block - 1) / warps_per_block);
#endif
}

extern "C" {

int hsp2g_abi_version(void) { return HSP2G_ABI_VERSION; }
const char *hsp2g_last_error(void) { return g_err.c_str(); }

const char *hsp2g_names(const char *t) {
    if (!t) return "";
    if (!strcmp(t, "params")) return k_params;
    if (!strcmp(t, "monthly")) return k_monthly;
    if (!strcmp(t, "flags")) return k_flags;
    if (!strcmp(t, "states")) return k_states;
    if (!strcmp(t, "forcings")) return k_forcings;
    if (!strcmp(t, "outputs")) return k_outputs;
    if (!strcmp(t, "errors")) return k_errors;
    if (!strcmp(t, "build")) {  // which build of the library this is: the product loader refuses anything but the CUDA one
#ifdef HSP2G_HOST_EMU
        return "host-emu";
#else
        return "cuda/sm_100a";
#endif
    }
    return "";
}

int hsp2g_device_count(int *count) {
    if (!count) return fail("null count");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        return fail("cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    return 0;
}

int hsp2g_batch_create(int device, int operation, int64_t n_seg, uint32_t act, double delt, hsp2g_batch **out) {
    if (!out) return fail("null out");
    *out = nullptr;
    if (operation != HSP2G_PERLND && operation != HSP2G_IMPLND) return fail("operation must be PERLND(0) or IMPLND(1)");
    if (n_seg <= 0 || n_seg > 2000000000LL) return fail("n_seg out of range");
    if (!(delt > 0.0) || delt > 1440.0) return fail("delt must be in (0, 1440] minutes");
    if (delt > 360.0 && (act & HSP2G_SNOW))
        return fail("SNOW with delt > 360: the reference raises KeyError('ICEFLG') (SNOW.py:85); refused");
    const unsigned perl = HSP2G_ATEMP | HSP2G_SNOW | HSP2G_PWATER | HSP2G_SEDMNT | HSP2G_PSTEMP | HSP2G_PWTGAS;
    const unsigned impl = HSP2G_ATEMP | HSP2G_SNOW | HSP2G_IWATER | HSP2G_SOLIDS | HSP2G_IWTGAS;
    const unsigned qual = operation == HSP2G_PERLND ? HSP2G_PQUAL : HSP2G_IQUAL;
    if (act != qual && (act == 0 || (act & ~(operation == HSP2G_PERLND ? perl : impl))))
        return fail("activity mask 0x%x is not valid for this operation (PQUAL / IQUAL rows form a batch of their own)", act);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail("no CUDA device (%s): libhsp2g has no CPU path", e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail("device %d out of range (%d devices)", device, ndev);
    DeviceGuard g(device);
    if (!g.ok) return fail("cudaSetDevice(%d) failed", device);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);

    hsp2g_batch *b = new hsp2g_batch();
    b->device = device;
    b->op = operation;
    b->n_seg = n_seg;
    b->ntiles = (n_seg + LAND_TILE - 1) / LAND_TILE;
    b->n_pad = b->ntiles * LAND_TILE;
    b->act = act;
    b->delt = delt;
    for (int i = 0; i < HSP2G_NMON; ++i) b->mslot[i] = -1;
    for (int i = 0; i < HSP2G_NFORC; ++i) {
        b->frow[i] = -1;
        b->ffill[i] = 0.0;
    }
    for (int i = 0; i < HSP2G_NOUT; ++i) b->srow[i] = -1;
    const size_t np = (size_t)b->n_pad;
    cudaError_t ce;
#define ALLOC(ptr, bytes)                                                          \
    if ((ce = cudaMalloc((void **)&(ptr), (bytes))) != cudaSuccess) {              \
        hsp2g_batch_destroy(b);                                                    \
        return fail("cudaMalloc(%zu bytes): %s", (size_t)(bytes), cudaGetErrorString(ce)); \
    }                                                                              \
    cudaMemset((ptr), 0, (bytes));
    ALLOC(b->segblk, np * SB_ROWS * sizeof(double));
    ALLOC(b->d_counter, 2 * sizeof(unsigned long long));
    ALLOC(b->d_progress, (size_t)b->ntiles * sizeof(unsigned));
#undef ALLOC
    ce = cudaStreamCreateWithFlags(&b->comp, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&b->h2d, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&b->d2h, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&b->last_launch, cudaEventDisableTiming);
    for (auto &s : b->slot) {
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&s.k_done, cudaEventDisableTiming);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&s.d2h_done, cudaEventDisableTiming);
    }
    if (ce != cudaSuccess) {
        hsp2g_batch_destroy(b);
        return fail("stream / event creation: %s", cudaGetErrorString(ce));
    }
    *out = b;
    return 0;
}

int hsp2g_batch_destroy(hsp2g_batch *b) {
    if (!b) return 0;
    DeviceGuard g(b->device);
    cudaDeviceSynchronize();
    cudaFree(b->segblk);
    cudaFree(b->mon);
    cudaFree(b->d_counter);
    cudaFree(b->d_progress);
    cudaFree(b->met);
    cudaFree(b->station);
    cudaFree(b->mfac);
    cudaFree(b->cal);
    cudaFree(b->d_periods);
    if (b->last_launch) cudaEventDestroy(b->last_launch);
    if (b->group_fork) cudaEventDestroy(b->group_fork);
    drop_kernel_image(b);
    for (auto &s : b->slot) {
        cudaFree(s.forc);
        cudaFree(s.out);
        cudaFree(s.agg);
        if (s.h2d_done) cudaEventDestroy(s.h2d_done);
        if (s.k_done) cudaEventDestroy(s.k_done);
        if (s.d2h_done) cudaEventDestroy(s.d2h_done);
    }
    for (auto &p : b->pending) {
        cudaEventDestroy(p.first);
        cudaEventDestroy(p.second);
    }
    for (auto e : b->ev_pool) cudaEventDestroy(e);
    if (b->comp) cudaStreamDestroy(b->comp);
    if (b->h2d) cudaStreamDestroy(b->h2d);
    if (b->d2h) cudaStreamDestroy(b->d2h);
    delete b;
    return 0;
}

int64_t hsp2g_batch_npad(const hsp2g_batch *b) { return b ? b->n_pad : 0; }

int hsp2g_set_params(hsp2g_batch *b, const double *par, int64_t ld) {
    if (!b || !par) return fail("null argument");
    if (ld < b->n_seg) return fail("ld < n_seg");
    DeviceGuard g(b->device);
    if (wait_last_launch(b)) return 1;  // a calibration loop re-parameterises between runs: never under a running kernel
    if (upload_rows(b->segblk, SB_ROWS, SB_PAR, par, ld, b->n_seg, b->ntiles, HSP2G_NPAR)) return 1;
    b->have_par = true;
    return 0;
}

int hsp2g_set_flags(hsp2g_batch *b, const uint64_t *flags) {
    if (!b || !flags) return fail("null argument");
    DeviceGuard g(b->device);
    if (wait_last_launch(b)) return 1;
    if (upload_rows(b->segblk, SB_ROWS, SB_FLAGS, flags, b->n_seg, b->n_seg, b->ntiles, 1)) return 1;
    b->flags_word = flags[0];
    b->flags_uniform = 1;
    b->flags_and = b->flags_or = flags[0];
    for (int64_t i = 1; i < b->n_seg; ++i) {
        b->flags_and &= flags[i];
        b->flags_or |= flags[i];
    }
    b->flags_uniform = b->flags_and == b->flags_or;
    b->have_flags = true;
    return 0;
}

int hsp2g_set_monthly(hsp2g_batch *b, int id, const double *tab, int64_t ld) {
    if (!b || !tab) return fail("null argument");
    if (id < 0 || id >= HSP2G_NMON) return fail("monthly table id %d out of range", id);
    if (ld < b->n_seg) return fail("ld < n_seg");
    DeviceGuard g(b->device);
    if (wait_last_launch(b)) return 1;
    const size_t one = (size_t)12 * b->n_seg;
    int slot = b->mslot[id];
    if (slot < 0) {  // a new table: grow the host copy and the tile-major device block by one slot
        slot = b->nmon_slots;
        double *grown = nullptr;
        CU(cudaDeviceSynchronize());
        CU(cudaMalloc((void **)&grown, (size_t)b->ntiles * 12 * (slot + 1) * LAND_TILE * sizeof(double)));
        cudaFree(b->mon);
        b->mon = grown;
        b->h_mon.resize(one * (slot + 1));
        // the slot is committed only now that its storage exists; its rows are uploaded below
        b->nmon_slots = slot + 1;
        b->mslot[id] = (short)slot;
    }
    for (int m = 0; m < 12; ++m)
        memcpy(&b->h_mon[one * slot + (size_t)m * b->n_seg], tab + (size_t)m * ld, (size_t)b->n_seg * sizeof(double));
    return upload_rows(b->mon, 12 * b->nmon_slots, 0, b->h_mon.data(), b->n_seg, b->n_seg, b->ntiles, 12 * b->nmon_slots);
}

int hsp2g_set_state(hsp2g_batch *b, const double *st, int64_t ld) {
    if (!b || !st) return fail("null argument");
    if (ld < b->n_seg) return fail("ld < n_seg");
    DeviceGuard g(b->device);
    if (wait_last_launch(b)) return 1;
    if (upload_rows(b->segblk, SB_ROWS, SB_STATE, st, ld, b->n_seg, b->ntiles, HSP2G_NSTATE)) return 1;
    b->have_state = true;
    return 0;
}

int hsp2g_get_state(hsp2g_batch *b, double *st, int64_t ld) {
    if (!b || !st) return fail("null argument");
    if (ld < b->n_seg) return fail("ld < n_seg");
    DeviceGuard g(b->device);
    if (wait_last_launch(b)) return 1;
    return download_rows(st, ld, b->segblk, SB_ROWS, SB_STATE, b->n_seg, b->ntiles, HSP2G_NSTATE);
}

int hsp2g_set_calendar(hsp2g_batch *b, int64_t n, const uint8_t *flags, const uint8_t *m0, const uint8_t *m1,
                       const double *xm, const double *dxm, const double *lapse) {
    if (!b || !flags || !m0 || !m1 || !xm || !dxm || !lapse) return fail("null argument");
    if (n <= 0) return fail("empty calendar");
    std::vector<CalStep> h((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        if (m0[i] > 11 || m1[i] > 11) return fail("calendar month index out of range at record %lld", (long long)i);
        CalStep c;
        memset(&c, 0, sizeof c);
        c.xm = xm[i];
        c.dxm = dxm[i];
        c.lapse = lapse[i];
        c.flags = flags[i];
        c.m0 = m0[i];
        c.m1 = m1[i];
        h[(size_t)i] = c;
    }
    DeviceGuard g(b->device);
    CU(cudaDeviceSynchronize());
    cudaFree(b->cal);
    b->cal = nullptr;
    // one zeroed record past the end: the kernels load the flags of interval t+1 while they work on t
    CU(cudaMalloc((void **)&b->cal, (size_t)(n + 1) * sizeof(CalStep)));
    CU(cudaMemset(b->cal, 0, (size_t)(n + 1) * sizeof(CalStep)));
    CU(cudaMemcpy(b->cal, h.data(), (size_t)n * sizeof(CalStep), cudaMemcpyHostToDevice));
    H2D_DONE();
    b->ncal = n;
    return 0;
}

int hsp2g_set_forcing_layout(hsp2g_batch *b, int n_rows, const int32_t *ids, const double *fill, uint32_t opts) {
    if (!b || !fill || (n_rows > 0 && !ids)) return fail("null argument");
    if (n_rows < 0 || n_rows > HSP2G_NFORC) return fail("n_rows out of range");
    short row[HSP2G_NFORC];
    for (int i = 0; i < HSP2G_NFORC; ++i) row[i] = -1;
    for (int r = 0; r < n_rows; ++r) {
        if (ids[r] < 0 || ids[r] >= HSP2G_NFORC) return fail("forcing id %d out of range", ids[r]);
        if (row[ids[r]] >= 0) return fail("forcing id %d listed twice", ids[r]);
        row[ids[r]] = (short)r;
    }
    if (b->met && n_rows != b->nf)
        return fail("the shared met table holds %d rows per interval: switch shared-met mode off (hsp2g_set_shared_forcing with "
                    "n_stations = 0) before changing the number of forcing rows", b->nf);
    memcpy(b->frow, row, sizeof row);
    memcpy(b->ffill, fill, sizeof(double) * HSP2G_NFORC);
    b->nf = n_rows;
    b->opts = opts;
    return 0;
}

int hsp2g_set_shared_forcing(hsp2g_batch *b, int32_t n_stations, int64_t n_steps, const double *series,
                             const int32_t *station, const double *mfactor, int64_t ld) {
    if (!b) return fail("null batch");
    DeviceGuard g(b->device);
    CU(cudaDeviceSynchronize());
    cudaFree(b->met);
    cudaFree(b->station);
    cudaFree(b->mfac);
    b->met = nullptr;
    b->station = nullptr;
    b->mfac = nullptr;
    b->n_stations = 0;
    b->n_met_steps = 0;
    if (n_stations == 0) return 0;  // back to per-segment forcings
    if (b->nf <= 0) return fail("hsp2g_set_forcing_layout must name the rows before hsp2g_set_shared_forcing");
    if (!series || !station || !mfactor) return fail("null argument");
    if (n_stations < 0 || n_steps <= 0 || ld < b->n_seg) return fail("bad shared forcing shape");
    std::vector<int> st((size_t)b->n_pad);
    for (int64_t i = 0; i < b->n_pad; ++i) {
        const int64_t seg = i < b->n_seg ? i : b->n_seg - 1;
        if (station[seg] < 0 || station[seg] >= n_stations) return fail("station index %d of segment %lld out of range", station[seg], (long long)seg);
        st[(size_t)i] = station[seg];
    }
    const size_t met_bytes = (size_t)n_steps * b->nf * n_stations * sizeof(double);
    CU(cudaMalloc((void **)&b->met, met_bytes));
    CU(cudaMemcpy(b->met, series, met_bytes, cudaMemcpyHostToDevice));
    CU(cudaMalloc((void **)&b->station, st.size() * sizeof(int)));
    CU(cudaMemcpy(b->station, st.data(), st.size() * sizeof(int), cudaMemcpyHostToDevice));
    H2D_DONE();
    CU(cudaMalloc((void **)&b->mfac, (size_t)b->ntiles * b->nf * LAND_TILE * sizeof(double)));
    if (upload_rows(b->mfac, b->nf, 0, mfactor, ld, b->n_seg, b->ntiles, b->nf)) return 1;
    b->n_stations = n_stations;
    b->n_met_steps = n_steps;
    return 0;
}

int hsp2g_set_save(hsp2g_batch *b, int n_rows, const int32_t *ids) {
    if (!b || (n_rows > 0 && !ids)) return fail("null argument");
    if (n_rows < 0 || n_rows > HSP2G_NOUT) return fail("n_rows out of range");
    short row[HSP2G_NOUT];
    for (int i = 0; i < HSP2G_NOUT; ++i) row[i] = -1;
    for (int r = 0; r < n_rows; ++r) {
        if (ids[r] < 0 || ids[r] >= HSP2G_NOUT) return fail("output id %d out of range", ids[r]);
        if (row[ids[r]] >= 0) return fail("output id %d listed twice", ids[r]);
        row[ids[r]] = (short)r;
    }
    memcpy(b->srow, row, sizeof row);
    b->ns = n_rows;
    return 0;
}

int hsp2g_set_series_layout(hsp2g_batch *b, int layout) {
    if (!b) return fail("null batch");
    if (layout != HSP2G_TILED && layout != HSP2G_PLAIN) return fail("layout must be HSP2G_TILED or HSP2G_PLAIN");
    b->plain = layout == HSP2G_PLAIN;
    return 0;
}

int hsp2g_set_forcing_dtype(hsp2g_batch *b, int dtype) {
    if (!b) return fail("null batch");
    if (dtype != HSP2G_F64 && dtype != HSP2G_F32) return fail("forcing dtype must be HSP2G_F64 or HSP2G_F32");
    b->forc_f32 = dtype == HSP2G_F32;
    return 0;
}

int hsp2g_run_device_ld(hsp2g_batch *b, int64_t t0, int32_t nsteps, const void *d_forc, int64_t ldf, void *d_out, int64_t ldo,
                        void *stream) {
    if (check_ready(b, t0, nsteps)) return 1;
    if (b->nf > 0 && !d_forc && !b->met) return fail("null forcing pointer");
    if (b->ns > 0 && !d_out) return fail("null output pointer");
    if (((uintptr_t)d_forc | (uintptr_t)d_out) & 15) return fail("device series must be 16-byte aligned (TMA bulk copies)");
    if (b->plain) {
        if ((b->nf > 0 && !b->met && ldf < b->n_pad) || (b->ns > 0 && ldo < b->n_pad))
            return fail("PLAIN device series need rows of at least %lld elements (32 x tiles)", (long long)b->n_pad);
        if ((ldf & 3) || (ldo & 3)) return fail("PLAIN row lengths must be multiples of 4 elements (16-byte rows)");
    }
    DeviceGuard g(b->device);
    return launch(b, t0, nsteps, d_forc, ldf, d_out, ldo, stream ? (cudaStream_t)stream : b->comp);
}

int hsp2g_run_device_linked(hsp2g_batch *b, int64_t t0, int32_t nsteps, int32_t n_rows, const hsp2g_row_source *rows,
                            int64_t src_tiles, void *d_out, int64_t ldo, void *stream) {
    if (check_ready(b, t0, nsteps)) return 1;
    if (b->met) return fail("a batch in shared-met mode has no per-segment forcing rows to link");
    if (b->forc_f32) return fail("linked forcing rows are float64 (the series another batch saved or was given)");
    if (n_rows != b->nf || n_rows > LAND_MAX_LINKED || (n_rows > 0 && !rows))
        return fail("%d row sources for a batch with %d forcing rows (at most %d can be linked)", n_rows, b->nf, LAND_MAX_LINKED);
    if (src_tiles <= 0 || src_tiles > b->ntiles) return fail("src_tiles must be in [1, %lld]", (long long)b->ntiles);
    for (int r = 0; r < n_rows; ++r)
        if (!rows[r].base || (((uintptr_t)rows[r].base | (uintptr_t)rows[r].interval_stride | (uintptr_t)rows[r].tile_stride) & 15))
            return fail("row source %d: base and strides must be non-null / multiples of 16 bytes (TMA bulk copies)", r);
    if (b->ns > 0 && !d_out) return fail("null output pointer");
    if ((uintptr_t)d_out & 15) return fail("device series must be 16-byte aligned (TMA bulk copies)");
    if (b->plain && b->ns > 0 && (ldo < b->n_pad || (ldo & 3)))
        return fail("PLAIN device series need rows of at least %lld elements, a multiple of 4", (long long)b->n_pad);
    DeviceGuard g(b->device);
    return launch(b, t0, nsteps, nullptr, b->n_pad, d_out, ldo, stream ? (cudaStream_t)stream : b->comp, rows, src_tiles);
}

int hsp2g_run_device(hsp2g_batch *b, int64_t t0, int32_t nsteps, const double *d_forc, double *d_out, void *stream) {
    if (!b) return fail("null batch");
    return hsp2g_run_device_ld(b, t0, nsteps, d_forc, b->n_pad, d_out, b->n_pad, stream);
}

void *hsp2g_batch_stream(hsp2g_batch *b) { return b ? (void *)b->comp : nullptr; }

int hsp2g_run_device_group(int32_t n_batches, hsp2g_batch *const *batches, int64_t t0, int32_t nsteps, const void *d_forc,
                           int64_t ldf, void *d_out, int64_t ldo, const int64_t *first_column, void *stream) {
    if (n_batches <= 0 || !batches || !first_column) return fail("bad group arguments");
    hsp2g_batch *b0 = batches[0];
    if (!b0) return fail("null batch");
    for (int k = 0; k < n_batches; ++k) {
        hsp2g_batch *b = batches[k];
        if (!b) return fail("null batch");
        if (b->device != b0->device || b->plain != b0->plain || b->nf != b0->nf || b->ns != b0->ns || b->out_f32 != b0->out_f32 ||
            b->forc_f32 != b0->forc_f32 || (b->met != nullptr) != (b0->met != nullptr))
            return fail("the batches of a group share one series buffer: same device, layout, row sets and element types");
        if (first_column[k] < 0 || (first_column[k] % LAND_TILE)) return fail("first_column must be a multiple of 32");
        if (check_ready(b, t0, nsteps)) return 1;
    }
    if (b0->nf > 0 && !d_forc && !b0->met) return fail("null forcing pointer");
    if (b0->ns > 0 && !d_out) return fail("null output pointer");
    DeviceGuard g(b0->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : b0->comp;
    const size_t fesz = b0->forc_f32 ? 4 : 8, osz = b0->out_f32 ? 4 : 8;
    // fork / join.  The groups run on at most GROUP_STREAMS of the batches' own streams (several groups of one stream run one
    // after the other), the largest groups first: more streams than the device has hardware queues (8 by default, up to 32 with
    // CUDA_DEVICE_MAX_CONNECTIONS, which hspsquared_b200/__init__.py raises) alias into the same queue, where one stream's
    // event record holds up the other's kernel -- measured: 48 streams ran 3-5x slower than one mixed kernel.  All launches
    // are enqueued before the caller's stream is made to wait for any of them, for the same reason.
    enum { GROUP_STREAMS = 24 };
    const int ns_used = n_batches < GROUP_STREAMS ? n_batches : GROUP_STREAMS;
    std::vector<int> order((size_t)n_batches);
    for (int k = 0; k < n_batches; ++k) order[(size_t)k] = k;
    std::stable_sort(order.begin(), order.end(), [&](int a, int c) { return batches[a]->n_seg > batches[c]->n_seg; });
    if (!b0->group_fork) CU(cudaEventCreateWithFlags(&b0->group_fork, cudaEventDisableTiming));
    CU(cudaEventRecord(b0->group_fork, st));
    std::vector<hsp2g_batch *> last_on((size_t)ns_used, nullptr);
    for (int i = 0; i < n_batches; ++i) {
        const int k = order[(size_t)i];
        hsp2g_batch *b = batches[k];
        cudaStream_t run_on = batches[order[(size_t)(i % ns_used)]]->comp;
        const int64_t c0 = first_column[k];
        const char *f = (const char *)d_forc;
        char *o = (char *)d_out;
        if (b->plain) {
            if (c0 + b->n_pad > ldf && b->nf > 0 && !b->met) return fail("group %d: columns beyond ld_forcing", k);
            if (c0 + b->n_pad > ldo && b->ns > 0) return fail("group %d: columns beyond ld_out", k);
            if (f) f += (size_t)c0 * fesz;
            if (o) o += (size_t)c0 * osz;
        } else {  // TILED: the group's tiles are contiguous
            if (f) f += (size_t)(c0 / LAND_TILE) * nsteps * b->nf * LAND_TILE * fesz;
            if (o) o += (size_t)(c0 / LAND_TILE) * nsteps * b->ns * LAND_TILE * osz;
        }
        if (i < ns_used && run_on != st) CU(cudaStreamWaitEvent(run_on, b0->group_fork, 0));
        if (launch(b, t0, nsteps, f, ldf, o, ldo, run_on)) return 1;
        last_on[(size_t)(i % ns_used)] = b;
    }
    for (int i = 0; i < ns_used; ++i)
        if (last_on[(size_t)i] && batches[order[(size_t)i]]->comp != st) {
            if (ensure_last_recorded(last_on[(size_t)i])) return 1;
            CU(cudaStreamWaitEvent(st, last_on[(size_t)i]->last_launch, 0));
        }
    return 0;
}

int hsp2g_run_host_ld(hsp2g_batch *b, int64_t t0, int32_t nsteps, const void *h_forc, int64_t ldf, void *h_out, int64_t ldo) {
    if (check_ready(b, t0, nsteps)) return 1;
    if (b->nf > 0 && !h_forc && !b->met) return fail("null forcing buffer");
    if (b->met) h_forc = nullptr;  // shared-met mode: nothing to copy in
    if (b->ns > 0 && !h_out) return fail("null output buffer");
    if (b->plain && ((h_forc && ldf < b->n_seg) || (b->ns > 0 && ldo < b->n_seg))) return fail("ld < n_seg");
    DeviceGuard g(b->device);
    Slot &s = b->slot[b->next_slot];
    b->next_slot ^= 1;
    const size_t fesz = b->forc_f32 ? 4 : 8, osz = b->out_f32 ? 4 : 8;
    const size_t frows = (size_t)nsteps * (size_t)(b->nf > 0 ? b->nf : 0), orows = (size_t)nsteps * (size_t)(b->ns > 0 ? b->ns : 0);
    const size_t need_f = h_forc ? (size_t)b->n_pad * fesz * frows : 0;
    const size_t need_o = (size_t)b->n_pad * osz * orows;
    if (need_f > s.cap_f || need_o > s.cap_o) {
        CU(cudaEventSynchronize(s.d2h_done));
        if (need_f > s.cap_f) {
            cudaFree(s.forc);
            s.forc = nullptr;
            s.cap_f = 0;
            CU(cudaMalloc((void **)&s.forc, need_f));
            // PLAIN: columns past n_seg of the last tile are never copied in; they read 0.  On the copy-in stream: the batch's
            // streams do not synchronise with the default stream, where a plain cudaMemset of device memory runs asynchronously
            // and could land AFTER the chunk's copy-in (seen on small batches: a group's forcings zeroed under its kernel)
            CU(cudaMemsetAsync(s.forc, 0, need_f, b->h2d));
            s.cap_f = need_f;
        }
        if (need_o > s.cap_o) {
            cudaFree(s.out);
            s.out = nullptr;
            s.cap_o = 0;
            CU(cudaMalloc((void **)&s.out, need_o));
            s.cap_o = need_o;
        }
    }
    // period statistics instead of series (hsp2g_set_output_periods): the chunk must be made of whole periods
    int nb = 0;
    size_t need_a = 0;
    if (!b->periods.empty() && b->ns > 0) {
        const std::vector<int> &P = b->periods;
        if (!b->out_f32) return fail("period statistics are defined on float32 outputs: hsp2g_set_output_dtype(HSP2G_F32)");
        if (t0 + nsteps > (int64_t)P.size()) return fail("intervals beyond the %lld given to hsp2g_set_output_periods", (long long)P.size());
        if ((t0 > 0 && P[t0 - 1] == P[t0]) || (t0 + nsteps < (int64_t)P.size() && P[t0 + nsteps] == P[t0 + nsteps - 1]))
            return fail("a chunk must start and end on period boundaries when period statistics are requested");
        nb = P[t0 + nsteps - 1] - P[t0] + 1;
        need_a = (size_t)nb * b->ns * 3 * b->n_pad * sizeof(float);
        if (need_a > s.cap_a) {
            CU(cudaEventSynchronize(s.d2h_done));
            cudaFree(s.agg);
            s.agg = nullptr;
            s.cap_a = 0;
            CU(cudaMalloc((void **)&s.agg, need_a));
            s.cap_a = need_a;
        }
    }
    // the slot is free once its previous outputs have left the device.  TILED: host and device share the tile-major layout, each
    // direction is ONE contiguous copy.  PLAIN: host rows are ld elements long, device rows n_pad: one strided (2-D) copy each
    // way, contiguous when the host rows are n_pad long too.
    CU(cudaStreamWaitEvent(b->h2d, s.d2h_done, 0));
    if (need_f) {
        if (!b->plain || ldf == b->n_pad)
            CU(cudaMemcpyAsync(s.forc, h_forc, need_f, cudaMemcpyHostToDevice, b->h2d));
        else
            CU(cudaMemcpy2DAsync(s.forc, (size_t)b->n_pad * fesz, h_forc, (size_t)ldf * fesz, (size_t)b->n_seg * fesz, frows,
                                 cudaMemcpyHostToDevice, b->h2d));
    }
    CU(cudaEventRecord(s.h2d_done, b->h2d));
    CU(cudaStreamWaitEvent(b->comp, s.h2d_done, 0));
    if (launch(b, t0, nsteps, s.forc, b->n_pad, s.out, b->n_pad, b->comp)) return 1;
    if (nb) {
        AggArgs A;
        A.src = (const float *)s.out;
        A.dst = s.agg;
#ifdef HSP2G_HOST_EMU
        A.bin = b->periods.data() + t0;
#else
        A.bin = b->d_periods + t0;
#endif
        A.bin_base = b->periods[(size_t)t0];
        A.plain = b->plain;
        A.ld = b->n_pad;
        A.n_threads = (long long)b->n_pad * b->ns;
        A.nsteps = nsteps;
        A.ns = b->ns;
        A.nbins = nb;
        if (run_aggregate(A, b->comp)) return 1;
    }
    CU(cudaEventRecord(s.k_done, b->comp));
    CU(cudaStreamWaitEvent(b->d2h, s.k_done, 0));
    const void *d_res = nb ? (const void *)s.agg : (const void *)s.out;
    const size_t res_rows = nb ? (size_t)nb * b->ns * 3 : orows, res_bytes = nb ? need_a : need_o;
    if (res_bytes) {
        if (!b->plain || ldo == b->n_pad)
            CU(cudaMemcpyAsync(h_out, d_res, res_bytes, cudaMemcpyDeviceToHost, b->d2h));
        else
            CU(cudaMemcpy2DAsync(h_out, (size_t)ldo * osz, d_res, (size_t)b->n_pad * osz, (size_t)b->n_seg * osz, res_rows,
                                 cudaMemcpyDeviceToHost, b->d2h));
    }
    CU(cudaEventRecord(s.d2h_done, b->d2h));
    return 0;
}

int hsp2g_run_host(hsp2g_batch *b, int64_t t0, int32_t nsteps, const double *h_forc, double *h_out) {
    if (!b) return fail("null batch");
    return hsp2g_run_host_ld(b, t0, nsteps, h_forc, b->plain ? b->n_seg : b->n_pad, h_out, b->plain ? b->n_seg : b->n_pad);
}

int hsp2g_pack_series(int64_t n_seg, int64_t nsteps, int32_t nrows, const double *src, int64_t ld, double *dst) {
    if (!src || !dst) return fail("null argument");
    if (n_seg <= 0 || nsteps < 0 || nrows < 0 || ld < n_seg) return fail("bad series shape");
    const int64_t ntiles = (n_seg + LAND_TILE - 1) / LAND_TILE;
    for (int64_t tl = 0; tl < ntiles; ++tl)
        for (int64_t t = 0; t < nsteps; ++t)
            for (int r = 0; r < nrows; ++r) {
                const double *in = src + ((size_t)t * nrows + r) * ld;
                double *o = dst + (((size_t)tl * nsteps + t) * nrows + r) * LAND_TILE;
                for (int l = 0; l < LAND_TILE; ++l) {
                    int64_t seg = tl * LAND_TILE + l;
                    o[l] = in[seg < n_seg ? seg : n_seg - 1];
                }
            }
    return 0;
}

int hsp2g_unpack_series(int64_t n_seg, int64_t nsteps, int32_t nrows, const double *src, double *dst, int64_t ld) {
    if (!src || !dst) return fail("null argument");
    if (n_seg <= 0 || nsteps < 0 || nrows < 0 || ld < n_seg) return fail("bad series shape");
    const int64_t ntiles = (n_seg + LAND_TILE - 1) / LAND_TILE;
    for (int64_t tl = 0; tl < ntiles; ++tl) {
        const int64_t n = n_seg - tl * LAND_TILE < LAND_TILE ? n_seg - tl * LAND_TILE : LAND_TILE;
        for (int64_t t = 0; t < nsteps; ++t)
            for (int r = 0; r < nrows; ++r)
                memcpy(dst + ((size_t)t * nrows + r) * ld + tl * LAND_TILE,
                       src + (((size_t)tl * nsteps + t) * nrows + r) * LAND_TILE, (size_t)n * sizeof(double));
    }
    return 0;
}

int hsp2g_sync(hsp2g_batch *b) {
    if (!b) return fail("null batch");
    DeviceGuard g(b->device);
    CU(cudaStreamSynchronize(b->h2d));
    CU(cudaStreamSynchronize(b->comp));
    CU(cudaStreamSynchronize(b->d2h));
    if (ensure_last_recorded(b)) return 1;
    if (b->have_last) CU(cudaEventSynchronize(b->last_launch));
    return drain_timing(b);
}

int hsp2g_sync_oldest(hsp2g_batch *b) {
    if (!b) return fail("null batch");
    DeviceGuard g(b->device);
    // the slot the NEXT hsp2g_run_host call will use holds the older of the (at most two) chunks in flight
    CU(cudaEventSynchronize(b->slot[b->next_slot].d2h_done));
    return 0;
}

int hsp2g_get_errors(hsp2g_batch *b, int64_t *err, int64_t ld) {
    if (!b || !err) return fail("null argument");
    if (ld < b->n_seg) return fail("ld < n_seg");
    DeviceGuard g(b->device);
    if (wait_last_launch(b)) return 1;
    return download_rows(err, ld, b->segblk, SB_ROWS, SB_ERR, b->n_seg, b->ntiles, HSP2G_NERR);
}

int hsp2g_reset_errors(hsp2g_batch *b) {
    if (!b) return fail("null batch");
    DeviceGuard g(b->device);
    if (wait_last_launch(b)) return 1;
    // on the compute stream, where the next launch goes (the default stream does not order itself with the batch's streams)
    CU(cudaMemset2DAsync(b->segblk + (size_t)SB_ERR * LAND_TILE, (size_t)SB_ROWS * LAND_TILE * 8, 0,
                         (size_t)HSP2G_NERR * LAND_TILE * 8, (size_t)b->ntiles, b->comp));
    CU(cudaStreamSynchronize(b->comp));
    return 0;
}

int hsp2g_math_probe(int device, int fn, int64_t n, const double *x, const double *y, double *out) {
    if (!x || !out || (fn == 0 && !y) || fn < 0 || fn > 3 || n <= 0) return fail("bad math probe arguments");
#ifdef HSP2G_HOST_EMU
    for (int64_t i = 0; i < n; ++i)
        out[i] = fn == 0 ? hsp_pow(x[i], y[i]) : fn == 1 ? hsp_exp(x[i]) : fn == 2 ? hsp_exp2(x[i]) : hsp_log(x[i]);
    (void)device;
    return 0;
#else
    DeviceGuard g(device);
    if (!g.ok) return fail("cudaSetDevice(%d) failed", device);
    double *dx = nullptr, *dy = nullptr, *dout = nullptr;
    const size_t bytes = (size_t)n * sizeof(double);
    CU(cudaMalloc((void **)&dx, bytes));
    CU(cudaMalloc((void **)&dy, bytes));
    CU(cudaMalloc((void **)&dout, bytes));
    CU(cudaMemcpy(dx, x, bytes, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dy, y ? y : x, bytes, cudaMemcpyHostToDevice));
    math_probe_kernel<<<(unsigned)((n + 255) / 256), 256>>>(fn, (long long)n, dx, dy, dout);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(out, dout, bytes, cudaMemcpyDeviceToHost);
    cudaFree(dx);
    cudaFree(dy);
    cudaFree(dout);
    if (e != cudaSuccess) return fail("math probe: %s", cudaGetErrorString(e));
    return 0;
#endif
}

struct hsp2g_gather {
    int device;
    long long n_dst, n_dst_pad, n_src;
    long long *seg_of;  // device [n_dst_pad]; padded entries repeat the last one
};

int hsp2g_gather_create(int device, int64_t n_src_seg, int64_t n_dst, const int64_t *seg_of, hsp2g_gather **out) {
    if (!out) return fail("null out");
    *out = nullptr;
    if (!seg_of || n_dst <= 0 || n_src_seg <= 0) return fail("bad gather arguments");
    for (int64_t j = 0; j < n_dst; ++j)
        if (seg_of[j] < 0 || seg_of[j] >= n_src_seg) return fail("seg_of[%lld] = %lld outside [0, %lld)", (long long)j, (long long)seg_of[j], (long long)n_src_seg);
    DeviceGuard g(device);
    if (!g.ok) return fail("cudaSetDevice(%d) failed", device);
    hsp2g_gather *h = new hsp2g_gather();
    h->device = device;
    h->n_dst = n_dst;
    h->n_src = n_src_seg;
    h->n_dst_pad = (n_dst + LAND_TILE - 1) / LAND_TILE * LAND_TILE;
    std::vector<long long> v((size_t)h->n_dst_pad);
    for (long long j = 0; j < h->n_dst_pad; ++j) v[(size_t)j] = seg_of[j < n_dst ? j : n_dst - 1];
    if (cudaMalloc((void **)&h->seg_of, v.size() * sizeof(long long)) != cudaSuccess) {
        delete h;
        return fail("out of device memory");
    }
    cudaMemcpy(h->seg_of, v.data(), v.size() * sizeof(long long), cudaMemcpyHostToDevice);
    cudaStreamSynchronize(0);  // see H2D_DONE
    *out = h;
    return 0;
}

void hsp2g_gather_destroy(hsp2g_gather *h) {
    if (!h) return;
    DeviceGuard g(h->device);
    cudaFree(h->seg_of);
    delete h;
}

int hsp2g_gather_run(hsp2g_gather *h, const double *d_src, int32_t src_rows, double *d_dst, int32_t dst_rows, int32_t k,
                     const int32_t *src_row, const int32_t *dst_row, int32_t nsteps, void *stream) {
    if (!h || !d_src || !d_dst || !src_row || !dst_row) return fail("null argument");
    if (k <= 0 || k > GATHER_MAXROWS) return fail("between 1 and %d rows per call", GATHER_MAXROWS);
    if (nsteps <= 0 || src_rows <= 0 || dst_rows <= 0) return fail("bad series shape");
    GatherArgs A;
    memset(&A, 0, sizeof A);
    A.src = d_src;
    A.dst = d_dst;
    A.seg_of = h->seg_of;
    A.n_dst_pad = h->n_dst_pad;
    A.nsteps = nsteps;
    A.src_rows = src_rows;
    A.dst_rows = dst_rows;
    A.k = k;
    for (int r = 0; r < k; ++r) {
        if (src_row[r] < 0 || src_row[r] >= src_rows || dst_row[r] < 0 || dst_row[r] >= dst_rows) return fail("row index out of range");
        A.src_row[r] = (short)src_row[r];
        A.dst_row[r] = (short)dst_row[r];
    }
    DeviceGuard g(h->device);
#ifdef HSP2G_HOST_EMU
    (void)stream;
    for (long long j = 0; j < A.n_dst_pad; ++j) {
        const long long seg = A.seg_of[j];
        for (int t = 0; t < nsteps; ++t)
            for (int r = 0; r < k; ++r)
                A.dst[(((size_t)(j / LAND_TILE) * nsteps + t) * dst_rows + A.dst_row[r]) * LAND_TILE + (j % LAND_TILE)] =
                    A.src[(((size_t)(seg / LAND_TILE) * nsteps + t) * src_rows + A.src_row[r]) * LAND_TILE + (seg % LAND_TILE)];
    }
    return 0;
#else
    gather_kernel<<<(unsigned)((A.n_dst_pad + 127) / 128), 128, 0, (cudaStream_t)stream>>>(A);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("gather launch failed: %s", cudaGetErrorString(e));
    return 0;
#endif
}

int hsp2g_aggregate_f32(int device, const float *d_src, int64_t ntiles, int32_t nsteps, int32_t ns, const int32_t *bin_of_step,
                        int32_t nbins, float *d_dst, void *stream) {
    if (!d_src || !d_dst || !bin_of_step) return fail("null argument");
    if (ntiles <= 0 || nsteps <= 0 || ns <= 0 || nbins <= 0) return fail("bad aggregation shape");
    for (int t = 0; t < nsteps; ++t)
        if (bin_of_step[t] < 0 || bin_of_step[t] >= nbins || (t && bin_of_step[t] < bin_of_step[t - 1]))
            return fail("bin_of_step must be non-decreasing within [0, nbins)");
    DeviceGuard g(device);
    if (!g.ok) return fail("cudaSetDevice(%d) failed", device);
    AggArgs A;
    A.src = d_src;
    A.dst = d_dst;
    A.bin_base = 0;
    A.plain = 0;
    A.ld = 0;
    A.n_threads = (long long)ntiles * ns * LAND_TILE;
    A.nsteps = nsteps;
    A.ns = ns;
    A.nbins = nbins;
#ifdef HSP2G_HOST_EMU
    A.bin = bin_of_step;
    return run_aggregate(A, (cudaStream_t)stream);
#else
    // stand-alone use: the bins travel to the device for this call (hsp2g_run_host keeps the whole run's periods resident)
    int *d_bin = nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaMalloc((void **)&d_bin, (size_t)nsteps * sizeof(int)));
    cudaError_t e = cudaMemcpyAsync(d_bin, bin_of_step, (size_t)nsteps * sizeof(int), cudaMemcpyHostToDevice, st);
    int rc = 0;
    if (e != cudaSuccess)
        rc = fail("cudaMemcpyAsync: %s", cudaGetErrorString(e));
    else {
        A.bin = d_bin;
        rc = run_aggregate(A, st);
    }
    cudaStreamSynchronize(st);
    cudaFree(d_bin);
    return rc;
#endif
}

int hsp2g_set_output_periods(hsp2g_batch *b, int64_t n, const int32_t *period_of_step) {
    if (!b) return fail("null batch");
    DeviceGuard g(b->device);
    if (n == 0 || !period_of_step) {  // back to series
        CU(cudaDeviceSynchronize());
        b->periods.clear();
        cudaFree(b->d_periods);
        b->d_periods = nullptr;
        return 0;
    }
    if (n < 0) return fail("n must be >= 0");
    for (int64_t t = 0; t < n; ++t)
        if (period_of_step[t] < 0 || (t && (period_of_step[t] < period_of_step[t - 1] || period_of_step[t] > period_of_step[t - 1] + 1)))
            return fail("period_of_step must start at >= 0 and grow by 0 or 1 from interval to interval");
    // the whole run's periods live on the device: a chunk's aggregation reads them in place, no per-chunk host copy
    CU(cudaDeviceSynchronize());
    cudaFree(b->d_periods);
    b->d_periods = nullptr;
    CU(cudaMalloc((void **)&b->d_periods, (size_t)n * sizeof(int)));
    CU(cudaMemcpy(b->d_periods, period_of_step, (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
    H2D_DONE();
    b->periods.assign(period_of_step, period_of_step + n);
    return 0;
}

// ---- kernels built for one batch (hspsquared_b200/jit.py).  The image is a cubin for sm_100a whose entry point `entry` takes
// the launch record by value and whose global `<entry>_info` (16 x u64) states what was fixed at compile time; launch() uses the
// kernel only while that still matches the batch (jit_matches).
int hsp2g_set_kernel_image(hsp2g_batch *b, const void *image, size_t nbytes, const char *entry, const char *display_name) {
    if (!b) return fail("null batch");
    DeviceGuard g(b->device);
    CU(cudaDeviceSynchronize());
    drop_kernel_image(b);
    if (!image || !nbytes) return 0;  // back to the library's own kernels
    if (!entry) return fail("null entry point name");
    const std::string info_name = std::string(entry) + "_info";
    unsigned long long info[JI_N];
#ifdef HSP2G_HOST_EMU
    // test double: `image` is the path of a host-compiled shared object with the same entry point
    const std::string path((const char *)image, nbytes);
    b->jit.dl = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!b->jit.dl) return fail("dlopen(%s): %s", path.c_str(), dlerror());
    b->jit.fn = (void (*)(const LandLaunch &))dlsym(b->jit.dl, entry);
    const void *ip = dlsym(b->jit.dl, info_name.c_str());
    if (!b->jit.fn || !ip) {
        drop_kernel_image(b);
        return fail("%s or %s missing from %s", entry, info_name.c_str(), path.c_str());
    }
    memcpy(info, ip, sizeof info);
    b->jit.resident = 1;
#else
    const Driver &d = driver();
    if (!d.ok) return fail("%s", d.why.c_str());
    CUresult r = d.ModuleLoadData(&b->jit.mod, image);
    if (r != CUDA_SUCCESS) return fail("cuModuleLoadData: %s", cu_str(r));
    CUdeviceptr ip = 0;
    size_t isz = 0;
    if ((r = d.ModuleGetFunction(&b->jit.fn, b->jit.mod, entry)) != CUDA_SUCCESS ||
        (r = d.ModuleGetGlobal(&ip, &isz, b->jit.mod, info_name.c_str())) != CUDA_SUCCESS || isz != sizeof info) {
        drop_kernel_image(b);
        return fail("kernel image lacks %s / %s: %s", entry, info_name.c_str(), cu_str(r));
    }
    CU(cudaMemcpy(info, (const void *)ip, sizeof info, cudaMemcpyDeviceToHost));
#endif
    if (info[JI_MAGIC] != HSP2G_JIT_MAGIC || info[JI_LAUNCH_BYTES] != sizeof(LandLaunch) || info[JI_SBROWS] != SB_ROWS ||
        info[JI_BLOCK] < LAND_TILE || info[JI_BLOCK] > 1024 || info[JI_BLOCK] % LAND_TILE) {
        drop_kernel_image(b);
        return fail("kernel image was built against other kernel sources than this library (launch record %llu vs %zu bytes)",
                    info[JI_LAUNCH_BYTES], sizeof(LandLaunch));
    }
#ifndef HSP2G_HOST_EMU
    if (info[JI_SMEM] > 48 * 1024 &&
        (r = driver().FuncSetAttribute(b->jit.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)info[JI_SMEM])) != CUDA_SUCCESS) {
        drop_kernel_image(b);
        return fail("cuFuncSetAttribute(max dynamic shared memory = %llu): %s", info[JI_SMEM], cu_str(r));
    }
    int per_sm = 0, sms = 0;
    r = driver().OccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, b->jit.fn, (int)info[JI_BLOCK], (size_t)info[JI_SMEM]);
    if (r != CUDA_SUCCESS || per_sm <= 0) {
        drop_kernel_image(b);
        return fail("the batch's kernel does not fit an SM (%llu bytes of shared memory): %s", info[JI_SMEM], cu_str(r));
    }
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, b->device));
    b->jit.resident = per_sm * sms;
#endif
    memcpy(b->jit.info, info, sizeof info);
    b->jit.name = display_name ? display_name : entry;
    b->jit.on = true;
    return 0;
}

int hsp2g_set_output_dtype(hsp2g_batch *b, int dtype) {
    if (!b) return fail("null batch");
    if (dtype != HSP2G_F64 && dtype != HSP2G_F32) return fail("output dtype must be HSP2G_F64 or HSP2G_F32");
    b->out_f32 = dtype == HSP2G_F32;
    return 0;
}

int hsp2g_set_schedule(hsp2g_batch *b, int32_t sub_steps) {
    if (!b) return fail("null batch");
    if (sub_steps < 0) return fail("sub_steps must be >= 0 (0 = one task per tile)");
    b->sub_steps = sub_steps;
    return 0;
}

int hsp2g_set_launch_overlap(hsp2g_batch *b, int on) {
    if (!b) return fail("null batch");
    DeviceGuard g(b->device);
    if (wait_last_launch(b)) return 1;  // the switch is thrown between launches, never under one
    b->overlap = on ? 1 : 0;
    return 0;
}

int64_t hsp2g_chained_launches(const hsp2g_batch *b) { return b ? b->chained_launches : 0; }

int hsp2g_set_timing(hsp2g_batch *b, int on) {
    if (!b) return fail("null batch");
    b->timing = on != 0;
    return 0;
}

int hsp2g_kernel_stats(hsp2g_batch *b, int64_t *launches, double *kernel_ms) {
    if (!b) return fail("null batch");
    DeviceGuard g(b->device);
    if (drain_timing(b)) return 1;
    if (launches) *launches = b->launches;
    if (kernel_ms) *kernel_ms = b->kernel_ms;
    return 0;
}

const char *hsp2g_kernel_name(const hsp2g_batch *b) { return b ? b->kname : ""; }

void *hsp2g_host_alloc_flags(size_t bytes, unsigned flags) {
    void *p = nullptr;
    unsigned f = cudaHostAllocDefault;
    if (flags & HSP2G_HOST_WRITE_COMBINED) f |= cudaHostAllocWriteCombined;
    if (flags & HSP2G_HOST_PORTABLE) f |= cudaHostAllocPortable;
    cudaError_t e = cudaHostAlloc(&p, bytes, f);
    if (e != cudaSuccess) {
        fail("cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
        return nullptr;
    }
    return p;
}
void *hsp2g_host_alloc(size_t bytes) { return hsp2g_host_alloc_flags(bytes, 0); }
int hsp2g_host_register(void *p, size_t bytes) {
    if (!p || !bytes) return fail("null argument");
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);  // one registration serves every device of the process
    if (e != cudaSuccess) {
        (void)cudaGetLastError();  // e.g. already page-locked: not an error of any later launch
        return fail("cudaHostRegister(%zu bytes): %s", bytes, cudaGetErrorString(e));
    }
    return 0;
}
int hsp2g_host_unregister(void *p) {
    if (p) {
        cudaError_t e = cudaHostUnregister(p);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            return fail("cudaHostUnregister: %s", cudaGetErrorString(e));
        }
    }
    return 0;
}
int hsp2g_host_free(void *p) {
    if (p) CU(cudaFreeHost(p));
    return 0;
}
void *hsp2g_device_alloc(int device, size_t bytes) {
    DeviceGuard g(device);
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        fail("cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
        return nullptr;
    }
    return p;
}
int hsp2g_device_free(int device, void *p) {
    DeviceGuard g(device);
    if (p) CU(cudaFree(p));
    return 0;
}
int hsp2g_memcpy_h2d(int device, void *dst, const void *src, size_t bytes) {
    DeviceGuard g(device);
    CU(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    H2D_DONE();
    return 0;
}
int hsp2g_memcpy_d2h(int device, void *dst, const void *src, size_t bytes) {
    DeviceGuard g(device);
    CU(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return 0;
}
}  // extern "C"
