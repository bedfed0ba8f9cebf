// perlnd_kernel.cu -- PERLND kernel instantiations and their launcher.
#include "land_kernels.cuh"
#include "launch.h"

template <int ACT_CT, bool HOURLY>
static cudaError_t launch_one(const LandLaunch &L, cudaStream_t st, const char **name, const char *nm) {
    const size_t smem = (size_t)PIPE_STAGES * L.nf * LAND_BLOCK * sizeof(double);
    auto k = perlnd_kernel<ACT_CT, HOURLY>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const unsigned grid = (unsigned)((L.n_seg + LAND_BLOCK - 1) / LAND_BLOCK);
    HSP2G_LAUNCH(k, grid, LAND_BLOCK, smem, st, L);
    if (name) *name = nm;
    return cudaGetLastError();
}

cudaError_t hsp2g_launch_perlnd(const LandLaunch &L, bool hourly, cudaStream_t st, const char **name) {
    const unsigned SP = HSP2G_ACT_SNOW | HSP2G_ACT_PWATER;
    if (L.act == SP)
        return hourly ? launch_one<(int)SP, true>(L, st, name, "perlnd_kernel<SNOW|PWATER,hourly>")
                      : launch_one<(int)SP, false>(L, st, name, "perlnd_kernel<SNOW|PWATER,subhourly>");
    return hourly ? launch_one<-1, true>(L, st, name, "perlnd_kernel<generic,hourly>")
                  : launch_one<-1, false>(L, st, name, "perlnd_kernel<generic,subhourly>");
}
