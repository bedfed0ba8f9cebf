// perlnd_kernel.cu -- PERLND kernel instantiations and their launcher.
#include "io_configs.h"
#include "land_kernels.cuh"
#include "launch.h"

#ifndef HSP_MAXREG_CHAIN
#define HSP_MAXREG_CHAIN 168  // full-chain kernels: 3 warps per SM sub-partition
#endif
#ifndef HSP_MAXREG_HOT
#define HSP_MAXREG_HOT 128  // register cap of the specialised kernels: 4 warps per SM sub-partition (16 per SM), no spills
#endif

// Kernels with a compile-time row set (StaticIO) are built for the TILED float64 layout; the general ones (DynIO) read the
// layout from the launch record (DynLay).  Kernels for any other exact configuration are built per batch (hspsquared_b200/jit.py).
template <int ACT_CT, bool HOURLY, class IO, int MAXREG, bool SHARED = false>
static cudaError_t launch_one(LandLaunch &L, int sub_pref, cudaStream_t st, const char **name, const char *nm, int *grid_out) {
    typedef typename LayFor<IO>::type LAY;
    const size_t smem = land_smem_bytes(IO::STATIC ? IO::NF : L.nf, perlnd_hot_io<IO>(ACT_CT), perlnd_ncold<HOURLY, IO, ACT_CT>(), SHARED,
                                        L.forc_f32 ? 4 : 8);
    auto k = perlnd_kernel<ACT_CT, HOURLY, IO, MAXREG, SHARED, DynFlags, LAY>;
    int cap = 0;
    cudaError_t e = hsp2g_prepare_kernel((const void *)k, smem, LAND_BLOCK, &cap);
    if (e != cudaSuccess) return e;
    hsp2g_plan_tasks(L, sub_pref, cap, LAND_WARPS, grid_out);
    HSP2G_LAUNCH(k, (unsigned)*grid_out, LAND_BLOCK, smem, st, L);
    if (name) *name = nm;
    return cudaGetLastError();
}

cudaError_t hsp2g_launch_perlnd(LandLaunch &L, bool hourly, int sub_pref, cudaStream_t st, const char **name, int *grid_out) {
    constexpr int SP = (int)(HSP2G_ACT_SNOW | HSP2G_ACT_PWATER);
    if (L.met) {  // shared-met mode: forcings gathered from the station table (hsp2g_set_shared_forcing)
        if (L.act == (unsigned)SP && hourly && io_matches<IO_SnowPwaterDefault>(L))
            return launch_one<SP, true, IO_SnowPwaterDefault, HSP_MAXREG_HOT, true>(L, sub_pref, st, name, "perlnd_kernel<SNOW|PWATER,hourly,save=default24,met=shared>", grid_out);
        if (L.act == (unsigned)SP && hourly && io_matches<IO_SnowPwaterDefaultF32>(L))
            return launch_one<SP, true, IO_SnowPwaterDefaultF32, HSP_MAXREG_HOT, true>(L, sub_pref, st, name, "perlnd_kernel<SNOW|PWATER,hourly,save=default24,f32,met=shared>", grid_out);
        return hourly ? launch_one<-1, true, DynIO, 232, true>(L, sub_pref, st, name, "perlnd_kernel<generic,hourly,io=dynamic,met=shared>", grid_out)
                      : launch_one<-1, false, DynIO, 232, true>(L, sub_pref, st, name, "perlnd_kernel<generic,subhourly,io=dynamic,met=shared>", grid_out);
    }
    if (L.act == (unsigned)SP && hourly) {
        if (io_matches<IO_SnowPwaterDefault>(L))
            return launch_one<SP, true, IO_SnowPwaterDefault, HSP_MAXREG_HOT>(L, sub_pref, st, name, "perlnd_kernel<SNOW|PWATER,hourly,save=default24>", grid_out);
        if (io_matches<IO_SnowPwaterDefaultF32>(L))
            return launch_one<SP, true, IO_SnowPwaterDefaultF32, HSP_MAXREG_HOT>(L, sub_pref, st, name, "perlnd_kernel<SNOW|PWATER,hourly,save=default24,f32>", grid_out);
        if (io_matches<IO_SnowPwaterFlows>(L))
            return launch_one<SP, true, IO_SnowPwaterFlows, HSP_MAXREG_HOT>(L, sub_pref, st, name, "perlnd_kernel<SNOW|PWATER,hourly,save=flows3>", grid_out);
        if (io_matches<IO_SnowPwaterAll>(L))
            return launch_one<SP, true, IO_SnowPwaterAll, 168>(L, sub_pref, st, name, "perlnd_kernel<SNOW|PWATER,hourly,save=all53>", grid_out);
    }
    constexpr int FULL = (int)(HSP2G_ACT_ATEMP | HSP2G_ACT_SNOW | HSP2G_ACT_PWATER | HSP2G_ACT_SEDMNT | HSP2G_ACT_PSTEMP |
                               HSP2G_ACT_PWTGAS);
    if (L.act == (unsigned)FULL && hourly) {
        if (io_matches<IO_FullChainAll>(L))
            return launch_one<FULL, true, IO_FullChainAll, HSP_MAXREG_CHAIN>(L, sub_pref, st, name, "perlnd_kernel<full chain,hourly,save=all84>", grid_out);
        if (io_matches<IO_FullChainCbp10>(L))
            return launch_one<FULL, true, IO_FullChainCbp10, HSP_MAXREG_CHAIN>(L, sub_pref, st, name, "perlnd_kernel<full chain,hourly,save=cbp10>", grid_out);
    }
    if (L.act == (unsigned)SP)
        return hourly ? launch_one<SP, true, DynIO, 168>(L, sub_pref, st, name, "perlnd_kernel<SNOW|PWATER,hourly,io=dynamic>", grid_out)
                      : launch_one<SP, false, DynIO, 168>(L, sub_pref, st, name, "perlnd_kernel<SNOW|PWATER,subhourly,io=dynamic>", grid_out);
    return hourly ? launch_one<-1, true, DynIO, 232>(L, sub_pref, st, name, "perlnd_kernel<generic,hourly,io=dynamic>", grid_out)
                  : launch_one<-1, false, DynIO, 232>(L, sub_pref, st, name, "perlnd_kernel<generic,subhourly,io=dynamic>", grid_out);
}
