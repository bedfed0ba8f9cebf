// implnd_kernel.cu -- IMPLND kernel instantiations and their launcher.
#include "io_configs.h"
#include "land_kernels.cuh"
#include "launch.h"

template <int ACT_CT, bool HOURLY, class IO, int MINB>
static cudaError_t launch_one(const LandLaunch &L, cudaStream_t st, const char **name, const char *nm) {
    const size_t smem = land_smem_bytes(IO::STATIC ? IO::NF : L.nf);
    auto k = implnd_kernel<ACT_CT, HOURLY, IO, MINB>;
    cudaError_t e = hsp2g_prepare_kernel((const void *)k, smem);
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)((L.ntiles + LAND_WARPS - 1) / LAND_WARPS);
    HSP2G_LAUNCH(k, grid, LAND_BLOCK, smem, st, L);
    if (name) *name = nm;
    return cudaGetLastError();
}

cudaError_t hsp2g_launch_implnd(const LandLaunch &L, bool hourly, cudaStream_t st, const char **name) {
    constexpr int SI = (int)(HSP2G_ACT_SNOW | HSP2G_ACT_IWATER);
    if (L.act == (unsigned)SI && hourly && io_matches<IO_SnowIwaterDefault>(L))
        return launch_one<SI, true, IO_SnowIwaterDefault, 4>(L, st, name, "implnd_kernel<SNOW|IWATER,hourly,save=default16>");
    return hourly ? launch_one<-1, true, DynIO, 3>(L, st, name, "implnd_kernel<generic,hourly,io=dynamic>")
                  : launch_one<-1, false, DynIO, 3>(L, st, name, "implnd_kernel<generic,subhourly,io=dynamic>");
}
