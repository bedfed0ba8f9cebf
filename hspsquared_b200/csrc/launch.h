// launch.h -- internal launcher interface between the C ABI (hsp2g_api.cu) and the kernel translation units.
#pragma once
#include "cuda_rt.h"

struct LandLaunch;
cudaError_t hsp2g_launch_perlnd(const LandLaunch &L, bool hourly, cudaStream_t st, const char **name);
cudaError_t hsp2g_launch_implnd(const LandLaunch &L, bool hourly, cudaStream_t st, const char **name);
