// implnd_kernel.cu -- IMPLND kernel instantiations and their launcher.
#include "land_kernels.cuh"
#include "launch.h"

template <int ACT_CT, bool HOURLY>
static cudaError_t launch_one(const LandLaunch &L, cudaStream_t st, const char **name, const char *nm) {
    const size_t smem = (size_t)PIPE_STAGES * L.nf * LAND_BLOCK * sizeof(double);
    auto k = implnd_kernel<ACT_CT, HOURLY>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const unsigned grid = (unsigned)((L.n_seg + LAND_BLOCK - 1) / LAND_BLOCK);
    HSP2G_LAUNCH(k, grid, LAND_BLOCK, smem, st, L);
    if (name) *name = nm;
    return cudaGetLastError();
}

cudaError_t hsp2g_launch_implnd(const LandLaunch &L, bool hourly, cudaStream_t st, const char **name) {
    return hourly ? launch_one<-1, true>(L, st, name, "implnd_kernel<generic,hourly>")
                  : launch_one<-1, false>(L, st, name, "implnd_kernel<generic,subhourly>");
}
