// qual_kernel.cu -- PQUAL / IQUAL: one thread = one (segment, constituent) row, one warp = one tile of 32 rows.
// Same machinery as the land kernels (land_common.cuh): the rows' input series -- the flows and sediment / solids outflows
// the constituent is associated with -- arrive through the per-warp TMA ring, (tile, sub-chunk) tasks are handed out by
// the ticket scheduler, the storage is carried in registers and handed on through the segment block, saved series stream
// out tile-major.  A constituent needs ~40 fp64 operations per interval against 7 series read and up to 20 written, so the
// kernel is bound by HBM like the land kernels; no hot-parameter staging is needed (the day's parameters sit in registers).
#include "io_configs.h"
#include "land_kernels.cuh"
#include "launch.h"
#include "qual.cuh"

#ifndef HSP_MAXREG_QUAL
#define HSP_MAXREG_QUAL 128  // no spills; 96 (20 warps) spills 88-128 B and is 7 % slower on PQUAL
#endif

template <bool IMP, class IO, int MAXREG, class LAY = TiledLay>
__global__ void __launch_bounds__(LAND_BLOCK) HSP_MAXNREG(MAXREG) qual_kernel(const __grid_constant__ LandLaunch L) {
    qual_body<IMP, IO, LAY>(L);
}

template <bool IMP, class IO>
static cudaError_t launch_qual(LandLaunch &L, int sub_pref, cudaStream_t st, const char **name, const char *nm, int *grid_out) {
    const size_t smem = land_smem_bytes(IO::STATIC ? IO::NF : L.nf, 0ull, 0, false, L.forc_f32 ? 4 : 8);
    auto k = qual_kernel<IMP, IO, HSP_MAXREG_QUAL, typename LayFor<IO>::type>;
    int cap = 0;
    cudaError_t e = hsp2g_prepare_kernel((const void *)k, smem, LAND_BLOCK, &cap);
    if (e != cudaSuccess) return e;
    hsp2g_plan_tasks(L, sub_pref, cap, LAND_WARPS, grid_out);
    HSP2G_LAUNCH(k, (unsigned)*grid_out, LAND_BLOCK, smem, st, L);
    if (name) *name = nm;
    return cudaGetLastError();
}

cudaError_t hsp2g_launch_qual(LandLaunch &L, bool implnd, int sub_pref, cudaStream_t st, const char **name, int *grid_out) {
    if (implnd) {
        if (io_matches<IO_IqualAll>(L)) return launch_qual<true, IO_IqualAll>(L, sub_pref, st, name, "qual_kernel<IQUAL,in=3,save=all12>", grid_out);
        return launch_qual<true, DynIO>(L, sub_pref, st, name, "qual_kernel<IQUAL,io=dynamic>", grid_out);
    }
    if (io_matches<IO_PqualAll>(L)) return launch_qual<false, IO_PqualAll>(L, sub_pref, st, name, "qual_kernel<PQUAL,in=7,save=all20>", grid_out);
    return launch_qual<false, DynIO>(L, sub_pref, st, name, "qual_kernel<PQUAL,io=dynamic>", grid_out);
}
