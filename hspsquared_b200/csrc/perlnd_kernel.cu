// perlnd_kernel.cu -- PERLND kernel instantiations and their launcher.
#include "io_configs.h"
#include "land_kernels.cuh"
#include "launch.h"

template <int ACT_CT, bool HOURLY, class IO, int MINB>
static cudaError_t launch_one(const LandLaunch &L, cudaStream_t st, const char **name, const char *nm) {
    const size_t smem = land_smem_bytes(IO::STATIC ? IO::NF : L.nf);
    auto k = perlnd_kernel<ACT_CT, HOURLY, IO, MINB>;
    cudaError_t e = hsp2g_prepare_kernel((const void *)k, smem);
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)((L.ntiles + LAND_WARPS - 1) / LAND_WARPS);
    HSP2G_LAUNCH(k, grid, LAND_BLOCK, smem, st, L);
    if (name) *name = nm;
    return cudaGetLastError();
}

cudaError_t hsp2g_launch_perlnd(const LandLaunch &L, bool hourly, cudaStream_t st, const char **name) {
    constexpr int SP = (int)(HSP2G_ACT_SNOW | HSP2G_ACT_PWATER);
    if (L.act == (unsigned)SP && hourly) {
        if (io_matches<IO_SnowPwaterDefault>(L))
            return launch_one<SP, true, IO_SnowPwaterDefault, 4>(L, st, name, "perlnd_kernel<SNOW|PWATER,hourly,save=default24>");
        if (io_matches<IO_SnowPwaterFlows>(L))
            return launch_one<SP, true, IO_SnowPwaterFlows, 4>(L, st, name, "perlnd_kernel<SNOW|PWATER,hourly,save=flows3>");
        if (io_matches<IO_SnowPwaterAll>(L))
            return launch_one<SP, true, IO_SnowPwaterAll, 4>(L, st, name, "perlnd_kernel<SNOW|PWATER,hourly,save=all53>");
    }
    if (L.act == (unsigned)SP)
        return hourly ? launch_one<SP, true, DynIO, 3>(L, st, name, "perlnd_kernel<SNOW|PWATER,hourly,io=dynamic>")
                      : launch_one<SP, false, DynIO, 3>(L, st, name, "perlnd_kernel<SNOW|PWATER,subhourly,io=dynamic>");
    return hourly ? launch_one<-1, true, DynIO, 2>(L, st, name, "perlnd_kernel<generic,hourly,io=dynamic>")
                  : launch_one<-1, false, DynIO, 2>(L, st, name, "perlnd_kernel<generic,subhourly,io=dynamic>");
}
