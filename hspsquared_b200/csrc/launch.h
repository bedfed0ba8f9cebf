// launch.h -- internal launcher interface between the C ABI (hsp2g_api.cu) and the kernel translation units.
#pragma once
#include "cuda_rt.h"
#include <stddef.h>

struct LandLaunch;
cudaError_t hsp2g_launch_perlnd(const LandLaunch &L, bool hourly, cudaStream_t st, const char **name);
cudaError_t hsp2g_launch_implnd(const LandLaunch &L, bool hourly, cudaStream_t st, const char **name);
// opt in to the dynamic shared memory a kernel needs and size the L1/shared carve-out for it (hsp2g_api.cu)
cudaError_t hsp2g_prepare_kernel(const void *kernel, size_t smem_bytes);
