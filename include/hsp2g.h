/*
 * hsp2g.h -- C ABI of libhsp2g: the B200-native (sm_100a CUDA) implementation of HSP2's land-segment hot path.
 *
 * What it replaces.  In respec/HSPsquared every land segment is advanced by Numba loops called one segment and
 * one activity at a time:
 *     errors, errmsgs = function(io_manager, siminfo, ui, ts)          src/hsp2/hsp2/main.py:249-250
 * with `function` one of atemp/snow/pwater/sedmnt/pstemp/pwtgas (PERLND) or atemp/snow/iwater/solids/iwtgas (IMPLND) from the
 * registry at src/hsp2/hsp2/configuration.py:45-55, each of which ends in an @njit core:
 *     _atemp_  ATEMP.py:33    _snow_   SNOW.py:91     _pwater_ PWATER.py:102 (+ proute :696)
 *     _iwater_ IWATER.py:75   _sedmnt_ SEDMNT.py:52   _pstemp_ PSTEMP.py:62  _pwtgas_ PWTGAS.py:65
 *     _solids_ SOLIDS.py:53   _iwtgas_ IWTGAS.py:65
 * This library runs the same arithmetic for a BATCH of independent segments of one operation,
 * all active sections fused into one pass per interval, in time chunks.  The reference has no FFI of its own (it is
 * pure Python); the binding a maintainer adds is the ctypes stub of hspsquared_b200/_lib.py, see INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns host buffers, the library owns device memory behind the handle;
 *   - every function returns 0 on success, non-zero on failure with a message in hsp2g_last_error() (thread local);
 *   - there is NO CPU fallback: without a CUDA device hsp2g_batch_create fails;
 *   - results are bit-identical to the reference's Numba kernels on a glibc 2.39 / FMA x86-64 host (every fp64 operation is
 *     IEEE-exact and pow / exp / exp2 / log restate that libm, csrc/glibc_math_core.h);
 *   - identifier spaces (parameter / monthly-table / flag-bit / state / forcing / output / error ids) are the
 *     positions in the comma separated lists returned by hsp2g_names(); hspsquared_b200/csrc/hsp2g_ids.h is the
 *     single definition;
 *   - per-segment tables (parameters, flags, monthly tables, states, error counters) are passed as row-major host
 *     arrays [row][segment] with leading dimension `ld` >= n_seg (elements); the library re-tiles them;
 *   - TIME SERIES (forcings in, SAVEd series out) use the device's own layout on both sides of the bus, so that a
 *     chunk moves as one contiguous copy and a warp's interval is one contiguous span (a single 1-D TMA bulk copy):
 *         [tile][interval][row][32]      tile = segment / 32, lane = segment % 32, ntiles = ceil(n_seg / 32)
 *     i.e. time-major [interval][segment] blocked by 32-segment tiles.  hsp2g_pack_series / hsp2g_unpack_series
 *     convert from / to plain [interval][row][segment]; lanes past n_seg in the last tile are padding (they carry
 *     copies of the last segment on input and are ignored on output).
 * A handle is not thread-safe; use one handle per GPU per thread (the reference kernels hold the GIL anyway).
 */
#ifndef HSP2G_H
#define HSP2G_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HSP2G_ABI_VERSION 3

#define HSP2G_PERLND 0 /* configuration.py:48-50 */
#define HSP2G_IMPLND 1 /* configuration.py:51-52 */

/* activity mask bits, registry order */
#define HSP2G_ATEMP 1u
#define HSP2G_SNOW 2u
#define HSP2G_PWATER 4u
#define HSP2G_SEDMNT 8u
#define HSP2G_PSTEMP 16u
#define HSP2G_PWTGAS 32u
#define HSP2G_IWATER 64u
#define HSP2G_SOLIDS 128u /* IMPLND SOLIDS, SOLIDS.py:53 */
#define HSP2G_IWTGAS 256u /* IMPLND IWTGAS, IWTGAS.py:65 */
/* Quality constituents, _pqual_ PQUAL.py:124 / _iqual_ IQUAL.py:109.  A batch created with ONLY one of these bits holds one
 * row per (segment, constituent): the reference loops constituents inside a segment, they do not interact, so each is a row
 * of its own.  The rows read the segment's SURO IFWO AGWO PERO WSSD SCRSD PREC (PQUAL) / SURO SOSLD PREC (IQUAL) as
 * forcing series.  Combining these bits with the hydrology sections in one batch is refused. */
#define HSP2G_PQUAL 512u
#define HSP2G_IQUAL 1024u

/* hsp2g_set_forcing_layout opts */
#define HSP2G_OPT_SED_RAIN_IS_PREC 1u /* SEDMNT.py:86-89: no RAINF in ts -> RAIN = PREC */
/* siminfo['units'] == 2 (SNOW.py:114-160,569-585, PWATER.py:193-244,660-688, IWATER.py:86-131,276-283, SEDMNT.py:113-117,258-262,
 * PSTEMP.py:80-126,136-149,205-215, PWTGAS.py:90-94,167-174,310-350, SOLIDS.py:84-88,143-145, IWTGAS.py:80-82,111-114,160-182): the parameters and
 * states handed to hsp2g_set_params / hsp2g_set_state are ALREADY converted to the kernels' English units in the reference's
 * expression order (hspsquared_b200/segstore.py does it); with this bit the kernels apply the conversions that happen per
 * interval -- parameter series after monthly interpolation, the series one section hands the next (stored in metric units,
 * read back in English ones), every stored series.  All land sections; PQUAL / IQUAL have no unit handling in the reference
 * and none here (their inputs are read in the units they were stored in). */
#define HSP2G_OPT_METRIC 2u

typedef struct hsp2g_batch hsp2g_batch;

int hsp2g_abi_version(void);
const char *hsp2g_last_error(void);
/* table in {"params","monthly","flags","states","forcings","outputs","errors"}; comma separated ids in order.
 * Output ids are prefixed with their section, e.g. "SNOW_PACKF", "PWATER_SURO", "IWATER_SURO".
 * "build" names the build of the library: "cuda/sm_100a" for the product; the host loader refuses anything else. */
const char *hsp2g_names(const char *table);
int hsp2g_device_count(int *count);

/* One batch = n_seg segments of one operation sharing one activity mask and one time base.
 * Replaces the per-segment dispatch loop of main.py:91-275 for land operations. */
int hsp2g_batch_create(int device, int operation, int64_t n_seg, uint32_t activity_mask, double delt_minutes,
                       hsp2g_batch **out);
int hsp2g_batch_destroy(hsp2g_batch *b);
int64_t hsp2g_batch_npad(const hsp2g_batch *b);

/* Segment store (replaces make_numba_dict, utilities.py:58-79, and the per-step broadcast of constants,
 * SNOW.py:49-51 / PWATER.py:51-53): structure-of-arrays fp64 parameters, one 64-bit option word per segment. */
int hsp2g_set_params(hsp2g_batch *b, const double *par /*[n_params][ld]*/, int64_t ld);
int hsp2g_set_flags(hsp2g_batch *b, const uint64_t *flags /*[n_seg]*/);
/* 12 monthly values per segment for one table (utilities.py:205-211 initm / :186-202 dayval). */
int hsp2g_set_monthly(hsp2g_batch *b, int table_id, const double *tab /*[12][ld]*/, int64_t ld);
/* Full carried state, named and hidden (hsp2tools/restart.py:9-14 plus SURVEY.md A.2-A.5). */
int hsp2g_set_state(hsp2g_batch *b, const double *state /*[n_states][ld]*/, int64_t ld);
int hsp2g_get_state(hsp2g_batch *b, double *state /*[n_states][ld]*/, int64_t ld);

/* Shared time base (replaces hoursval/hourflag/monthval/dayval per-step arrays, utilities.py:136-202):
 * n = steps+1 records keyed by interval start.  flags = HRFG|DAYFG<<1|HR1FG<<2|HR6IND<<3|SEASONS<<4|NEWDAY<<5. */
int hsp2g_set_calendar(hsp2g_batch *b, int64_t n, const uint8_t *flags, const uint8_t *m0, const uint8_t *m1,
                       const double *xm, const double *dxm, const double *lapse);

/* Which forcing ids the forcing series carry (row order = ids order), and the value every absent id reads (the
 * reference wrappers' NaN / zero / -1e30 fills; get_timeseries utilities.py:274-290).  Rows in ascending id order
 * let the library pick a kernel specialised at compile time for the row set (see hsp2g_kernel_name). */
int hsp2g_set_forcing_layout(hsp2g_batch *b, int n_rows, const int32_t *forcing_ids,
                             const double *fill /*[n_forcings]*/, uint32_t opts);
/* Shared-met mode (replaces get_timeseries' per-segment `series * MFACTOR`, utilities.py:274-290, for watersheds whose
 * segments draw on a few met stations): the whole run's series of every forcing row, one column per station, live in
 * HBM once -- series[interval][row][station], rows as named by hsp2g_set_forcing_layout -- and interval t of segment i
 * reads series[t][row][station[i]] * mfactor[row][i] (one IEEE multiplication, what `temp1 *= MFACTOR` does to a
 * float64 series).  No forcing crosses PCIe per chunk and the kernels read no per-segment forcing from HBM:
 * hsp2g_run_device / hsp2g_run_host then take a NULL forcing pointer.  n_stations = 0 switches the mode off. */
int hsp2g_set_shared_forcing(hsp2g_batch *b, int32_t n_stations, int64_t n_steps, const double *series,
                             const int32_t *station /*[n_seg]*/, const double *mfactor /*[n_rows][ld]*/, int64_t ld);

/* Which output ids are written (the SAVE table, utilities.py:292-333), row order = ids order (ascending id order
 * enables the specialised kernels). */
int hsp2g_set_save(hsp2g_batch *b, int n_rows, const int32_t *output_ids);

/* Element type of the SAVEd series the kernels write (default HSP2G_F64).  HSP2G_F32 rounds each value to float32
 * (nearest even) in the kernel's store, which is exactly what save_timeseries hands the IOManager
 * (`df.astype(np.float32)`, utilities.py:313): half the bytes to HBM and over PCIe.  The series layout is unchanged,
 * [tile][interval][row][32] of the chosen type. */
#define HSP2G_F64 0
#define HSP2G_F32 1
int hsp2g_set_output_dtype(hsp2g_batch *b, int dtype);

/* Series layout of hsp2g_run_device / hsp2g_run_host (default HSP2G_TILED, described at the top of this file).
 * HSP2G_PLAIN is the orientation the reference itself keeps -- ts[name] is one array per series, a batch is
 * [interval][row][segment] (get_timeseries / save_timeseries, utilities.py:274-333): forcing [nsteps][n_forcing_rows][ld],
 * out [nsteps][n_saved_rows][ld].  Nothing is re-tiled on the host: a warp fetches its interval's box of n_forcing_rows x 32
 * segments with one tensor-map TMA copy and stores 256 contiguous bytes per saved row.  On the device ld >= 32*ntiles
 * (hsp2g_batch_npad); host buffers may have any ld >= n_seg. */
#define HSP2G_TILED 0
#define HSP2G_PLAIN 1
int hsp2g_set_series_layout(hsp2g_batch *b, int layout);
/* Element type of the forcing series (default HSP2G_F64).  HSP2G_F32: the series are float32 and are widened on the device --
 * exact whenever the source is float32 (WDM and the HDF5 /TIMESERIES tables are; `series * MFACTOR` stays in the stored
 * dtype, utilities.py:283-287) -- half the bytes over PCIe and out of HBM.  Ignored in shared-met mode. */
int hsp2g_set_forcing_dtype(hsp2g_batch *b, int dtype);

/* Advance all segments over intervals [t0, t0+nsteps).  State is read at the start and written at the end, so
 * consecutive chunks are bit-identical to one pass.  Series are tile-major [tile][nsteps][row][32] (see above):
 * forcing = ntiles*nsteps*n_forcing_rows*32 doubles, out = ntiles*nsteps*n_saved_rows*32 doubles (or floats).
 * _device: forcings and outputs already in HBM, 16-byte aligned; `stream` is a cudaStream_t (NULL = the batch's own).
 * _host:   host buffers; H2D, kernel and D2H are pipelined on three streams over two device slots; buffers must stay
 *          valid until hsp2g_sync.  Pinned host memory (hsp2g_host_alloc) is needed for real overlap. */
int hsp2g_run_device(hsp2g_batch *b, int64_t t0, int32_t nsteps, const double *d_forcing, double *d_out, void *stream);
int hsp2g_run_host(hsp2g_batch *b, int64_t t0, int32_t nsteps, const double *h_forcing, double *h_out);
/* The same with explicit row lengths (elements) for the HSP2G_PLAIN layout: forcing [nsteps][rows][ld_forcing] of the forcing
 * dtype, out [nsteps][rows][ld_out] of the output dtype (period statistics: [periods][rows][3][ld_out] float32).  _device:
 * ld >= hsp2g_batch_npad, a multiple of 4; _host: ld >= n_seg (one strided copy each way; contiguous when ld = npad).
 * hsp2g_run_device / hsp2g_run_host use ld = npad / ld = n_seg when the layout is PLAIN. */
int hsp2g_run_device_ld(hsp2g_batch *b, int64_t t0, int32_t nsteps, const void *d_forcing, int64_t ld_forcing, void *d_out,
                        int64_t ld_out, void *stream);
int hsp2g_run_host_ld(hsp2g_batch *b, int64_t t0, int32_t nsteps, const void *h_forcing, int64_t ld_forcing, void *h_out,
                      int64_t ld_out);
/* Forcing rows read IN PLACE from other device buffers -- how the PQUAL / IQUAL rows take SURO, IFWO, AGWO, PERO, WSSD, SCRSD and
 * PREC from the buffers the land batch of the same chunk wrote and read (the reference's constituents read the segment's ts
 * arrays in place too, PQUAL.py:119-126, IQUAL.py:104-109): no copy, no scratch buffer.  Forcing row r (ascending-id order, as for
 * hsp2g_run_device) of interval t of this batch's tile q is the 32 float64 at
 *     rows[r].base + t * interval_stride + (q % src_tiles) * tile_stride          (bytes; all multiples of 16)
 * -- for a source row `row` of a TILED buffer with R rows: base = buf + row*256, interval_stride = R*256, tile_stride =
 * nsteps*R*256; of a PLAIN buffer with row length ld: base = buf + row*ld*8, interval_stride = R*ld*8, tile_stride = 256.  A batch
 * of (segment, constituent) rows laid out constituent-major over the source's tiles (entity k*src_tiles*32 + p = constituent k of
 * the source's device segment p) reads every source row once per constituent.  n_rows = the batch's forcing rows (<= 16). */
typedef struct hsp2g_row_source {
    const void *base;
    int64_t interval_stride, tile_stride;
} hsp2g_row_source;
int hsp2g_run_device_linked(hsp2g_batch *b, int64_t t0, int32_t nsteps, int32_t n_rows, const hsp2g_row_source *rows,
                            int64_t src_tiles, void *d_out, int64_t ld_out, void *stream);
/* Several batches over ONE series buffer, launched so that they overlap on the device: batch k owns the columns (segments)
 * [first_column[k], first_column[k] + npad_k) of forcing / out (first_column a multiple of 32; PLAIN: ld >= the last column;
 * TILED: its tiles are contiguous).  This is how an ensemble with several option words runs: its segments are grouped by word
 * (divergence-aware sorting), each group is a batch with the kernel built for its word, and the groups -- each too small to
 * fill the GPU -- share it.  Every batch runs on its own stream, forked from and joined to `stream` (NULL = batches[0]'s). */
int hsp2g_run_device_group(int32_t n_batches, hsp2g_batch *const *batches, int64_t t0, int32_t nsteps, const void *d_forcing,
                           int64_t ld_forcing, void *d_out, int64_t ld_out, const int64_t *first_column, void *stream);
/* A kernel built for exactly this batch -- activity mask, forcing rows, SAVE set, element types, layout and, when every segment
 * carries the same option word, that word fixed at compile time (hspsquared_b200/jit.py generates the instantiation from
 * csrc/land_kernels.cuh and compiles it with `nvcc -cubin`).  image = the cubin (sm_100a), entry = its kernel, whose companion
 * global `<entry>_info` states what was fixed; the library launches it as long as that still describes the batch and runs its
 * own general kernels otherwise (hsp2g_kernel_name tells).  image = NULL drops the kernel. */
int hsp2g_set_kernel_image(hsp2g_batch *b, const void *image, size_t nbytes, const char *entry, const char *display_name);
/* Host-side re-tiling between [interval][row][ld] and [tile][interval][row][32] (what get_timeseries /
 * save_timeseries hand over per segment, utilities.py:274-333, for a whole batch). */
/* Period aggregation of saved float32 series on the device: what IOManager.write_ts computes with pandas for outstep 3 / 4 / 5
 * (hsp2io/io.py:74-94: last, sum and mean per day / month / year of the float32 frame save_timeseries hands it).
 * d_src [tile][nsteps][ns][32] float32 (hsp2g_set_output_dtype(HSP2G_F32)); bin_of_step[t] = period of interval t (host array,
 * non-decreasing, < nbins; hspsquared_b200/results.py period_bins computes it: days are 24-hour periods anchored at the run's
 * first label, as `resample('D', origin='start')` bins them in the pandas the reference pins, months / years calendar periods); d_dst
 * [tile][nbins][ns][3][32] float32 with the three statistics in the order last, sum, mean.  Sums are pandas' compensated
 * float32 sums over the non-NaN values; an all-NaN period gives last = NaN, sum = 0, mean = NaN.  One call covers whole periods:
 * cut the run into chunks that end where periods end (the compensated sum of a period is not carried from call to call). */
int hsp2g_aggregate_f32(int device, const float *d_src, int64_t ntiles, int32_t nsteps, int32_t ns, const int32_t *bin_of_step,
                        int32_t nbins, float *d_dst, void *stream);

/* Make hsp2g_run_host deliver period statistics instead of series (outstep 3 / 4 / 5 of main.py:259-262 -> io.py:74-94):
 * period_of_step[t] = period of interval t of the run (grows by 0 or 1), n = 0 switches back to series.  Needs float32 outputs.
 * Every chunk given to hsp2g_run_host must then consist of whole periods; its h_out receives
 * [tile][periods of the chunk][row][3 = last, sum, mean][32] float32 -- 3 / (intervals per period) of the series' bytes cross PCIe. */
int hsp2g_set_output_periods(hsp2g_batch *b, int64_t n, const int32_t *period_of_step);

/* The stream hsp2g_run_device uses when it is given none (non-blocking; nothing orders it against other streams): pass it to
 * hsp2g_gather_run, or to another batch's hsp2g_run_device, to chain work on the device in order. */
void *hsp2g_batch_stream(hsp2g_batch *b);

/* Row gather between two device-resident tile-major series buffers with the same interval count: destination entity j takes
 * row dst_row[r] from row src_row[r] of source segment seg_of[j] (r < k <= 16 per call).  Feeds a PQUAL / IQUAL batch, whose
 * entities are (segment, constituent) rows, from the land batch's output and forcing buffers without leaving the device; the
 * reference has no counterpart (its constituents read the segment's ts arrays in place, PQUAL.py:119-126). */
typedef struct hsp2g_gather hsp2g_gather;
int hsp2g_gather_create(int device, int64_t n_src_seg, int64_t n_dst, const int64_t *seg_of, hsp2g_gather **out);
int hsp2g_gather_run(hsp2g_gather *g, const double *d_src, int32_t src_rows, double *d_dst, int32_t dst_rows, int32_t k,
                     const int32_t *src_row, const int32_t *dst_row, int32_t nsteps, void *stream);
void hsp2g_gather_destroy(hsp2g_gather *g);

int hsp2g_pack_series(int64_t n_seg, int64_t nsteps, int32_t nrows, const double *src, int64_t ld, double *dst_tiled);
int hsp2g_unpack_series(int64_t n_seg, int64_t nsteps, int32_t nrows, const double *src_tiled, double *dst, int64_t ld);
int hsp2g_sync(hsp2g_batch *b);
/* Wait only until the OLDER of the (at most two) chunks in flight through hsp2g_run_host has arrived in its host buffer: a
 * caller that streams results out (save_timeseries' job, utilities.py:292-333) consumes chunk k-1 while chunk k computes. */
int hsp2g_sync_oldest(hsp2g_batch *b);

/* Error counters (the `errors` vectors returned beside ERRMSGS, main.py:255-257), int64 [n_errors][ld]. */
int hsp2g_get_errors(hsp2g_batch *b, int64_t *err, int64_t ld);
int hsp2g_reset_errors(hsp2g_batch *b);

/* Scheduling of one launch.  When a batch has more 32-segment tiles than the GPU holds warps, a launch is cut into
 * tasks of `sub_steps` intervals of one tile, drawn in order by the resident warps, so the device stays full to the
 * end of the chunk; a tile's state passes from task to task through the segment store, results are bit-identical.
 * Default 64; 0 = one task per tile (whole chunk). */
int hsp2g_set_schedule(hsp2g_batch *b, int32_t sub_steps);
/* Overlap of consecutive launches (off by default).  A launch ends with a tail: the last tasks finish one by one while the rest
 * of the device idles (4 % of a 438-interval launch over 100,000 segments).  With overlap on, an hsp2g_run_device /
 * hsp2g_run_device_ld call whose predecessor IN THE STREAM is the previous launch of the same batch is made as a programmatic
 * dependent launch: its blocks take the places the previous launch's blocks free as they run out of tasks, and every tile
 * waits for its own previous chunk only (the per-tile hand-over the tasks of one launch already use), so results are bit-identical.
 * The caller's side of the contract: the forcing buffer of a call is complete, and its output buffer free and distinct from the
 * previous call's, when the PREVIOUS call is enqueued -- anything enqueued on the stream between two calls (a copy, an event, a
 * kernel of another batch) orders the second call after it in the ordinary way and simply switches the overlap off for that pair.
 * Applies to batches that run their own kernel (hsp2g_set_kernel_image) on a grid that fills the device, to launches cut into
 * at least two tasks per tile. */
int hsp2g_set_launch_overlap(hsp2g_batch *b, int on);
/* launches made that way so far (instrumentation: tests check that the overlap really happened) */
int64_t hsp2g_chained_launches(const hsp2g_batch *b);

/* Instrumentation: kernels launched so far and, when timing is on, the summed CUDA-event duration of those kernels
 * (events on the launching stream). */
int hsp2g_set_timing(hsp2g_batch *b, int on);
int hsp2g_kernel_stats(hsp2g_batch *b, int64_t *launches, double *kernel_ms);
const char *hsp2g_kernel_name(const hsp2g_batch *b);

/* The kernels' pow / exp / exp2 / log restate glibc 2.39's libm -- the library the reference's Numba code reaches --
 * bit for bit (hspsquared_b200/csrc/glibc_math_core.h).  This probe evaluates them on the device for n arguments
 * (fn: 0 pow(x,y), 1 exp(x), 2 exp2(x), 3 log(x); y may be NULL unless fn == 0) so that a test can compare with libm. */
int hsp2g_math_probe(int device, int fn, int64_t n, const double *x, const double *y, double *out);

/* Pinned host memory / raw device memory helpers for callers without a CUDA runtime of their own. */
void *hsp2g_host_alloc(size_t bytes);
/* flags: HSP2G_HOST_WRITE_COMBINED for buffers the host only writes and the device reads (forcings: no cache snooping on the way
 * over PCIe); hsp2g_host_register pins an existing allocation (a numpy array, say) so that copies from / to it run asynchronously. */
#define HSP2G_HOST_WRITE_COMBINED 1u
#define HSP2G_HOST_PORTABLE 2u
void *hsp2g_host_alloc_flags(size_t bytes, unsigned flags);
int hsp2g_host_register(void *p, size_t bytes);
int hsp2g_host_unregister(void *p);
int hsp2g_host_free(void *p);
void *hsp2g_device_alloc(int device, size_t bytes);
int hsp2g_device_free(int device, void *p);
int hsp2g_memcpy_h2d(int device, void *dst, const void *src, size_t bytes);
int hsp2g_memcpy_d2h(int device, void *dst, const void *src, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif
