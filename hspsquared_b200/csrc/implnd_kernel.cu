// implnd_kernel.cu -- IMPLND kernel instantiations and their launcher.
#include "io_configs.h"
#include "land_kernels.cuh"
#include "launch.h"

template <int ACT_CT, bool HOURLY, class IO, int MAXREG, bool SHARED = false>
static cudaError_t launch_one(LandLaunch &L, int sub_pref, cudaStream_t st, const char **name, const char *nm, int *grid_out) {
    typedef typename LayFor<IO>::type LAY;
    const size_t smem = land_smem_bytes(IO::STATIC ? IO::NF : L.nf, IMPLND_HOT, implnd_ncold<HOURLY, IO>(), SHARED, L.forc_f32 ? 4 : 8);
    auto k = implnd_kernel<ACT_CT, HOURLY, IO, MAXREG, SHARED, DynFlags, LAY>;
    int cap = 0;
    cudaError_t e = hsp2g_prepare_kernel((const void *)k, smem, LAND_BLOCK, &cap);
    if (e != cudaSuccess) return e;
    hsp2g_plan_tasks(L, sub_pref, cap, LAND_WARPS, grid_out);
    HSP2G_LAUNCH(k, (unsigned)*grid_out, LAND_BLOCK, smem, st, L);
    if (name) *name = nm;
    return cudaGetLastError();
}

cudaError_t hsp2g_launch_implnd(LandLaunch &L, bool hourly, int sub_pref, cudaStream_t st, const char **name, int *grid_out) {
    constexpr int SI = (int)(HSP2G_ACT_SNOW | HSP2G_ACT_IWATER);
    if (L.met)
        return hourly ? launch_one<-1, true, DynIO, 128, true>(L, sub_pref, st, name, "implnd_kernel<generic,hourly,io=dynamic,met=shared>", grid_out)
                      : launch_one<-1, false, DynIO, 128, true>(L, sub_pref, st, name, "implnd_kernel<generic,subhourly,io=dynamic,met=shared>", grid_out);
    // test10's IMPLND chain: SNOW -> IWATER -> SOLIDS (-> IWTGAS); SaveTable.csv saves nothing of SOLIDS / IWTGAS by default
    constexpr int SIS = SI | (int)HSP2G_ACT_SOLIDS, SISG = SIS | (int)HSP2G_ACT_IWTGAS;
    if (hourly && io_matches<IO_SnowIwaterDefault>(L)) {
        if (L.act == (unsigned)SI)
            return launch_one<SI, true, IO_SnowIwaterDefault, 128>(L, sub_pref, st, name, "implnd_kernel<SNOW|IWATER,hourly,save=default16>", grid_out);
        if (L.act == (unsigned)SIS)
            return launch_one<SIS, true, IO_SnowIwaterDefault, 128>(L, sub_pref, st, name, "implnd_kernel<SNOW|IWATER|SOLIDS,hourly,save=default16>", grid_out);
        if (L.act == (unsigned)SISG)
            return launch_one<SISG, true, IO_SnowIwaterDefault, 128>(L, sub_pref, st, name, "implnd_kernel<SNOW|IWATER|SOLIDS|IWTGAS,hourly,save=default16>", grid_out);
    }
    return hourly ? launch_one<-1, true, DynIO, 128>(L, sub_pref, st, name, "implnd_kernel<generic,hourly,io=dynamic>", grid_out)
                  : launch_one<-1, false, DynIO, 128>(L, sub_pref, st, name, "implnd_kernel<generic,subhourly,io=dynamic>", grid_out);
}
