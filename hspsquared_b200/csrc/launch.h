// launch.h -- internal launcher interface between the C ABI (hsp2g_api.cu) and the kernel translation units.
#pragma once
#include "cuda_rt.h"
#include <stddef.h>

struct LandLaunch;
// Fill in the task plan of L (sub, nsub), launch, and report the grid size used (tickets drawn = tasks + grid).
cudaError_t hsp2g_launch_perlnd(LandLaunch &L, bool hourly, int sub_pref, cudaStream_t st, const char **name, int *grid_out);
cudaError_t hsp2g_launch_implnd(LandLaunch &L, bool hourly, int sub_pref, cudaStream_t st, const char **name, int *grid_out);
// PQUAL (implnd = false) / IQUAL rows: a batch whose activity mask is that one bit (qual_kernel.cu)
cudaError_t hsp2g_launch_qual(LandLaunch &L, bool implnd, int sub_pref, cudaStream_t st, const char **name, int *grid_out);
// opt in to the dynamic shared memory a kernel needs; *resident = blocks of `block` threads the device holds at once
cudaError_t hsp2g_prepare_kernel(const void *kernel, size_t smem_bytes, int block, int *resident);
// tasks = (tile, sub-chunk of sub_pref intervals) when the tiles outnumber the resident warps, else one task per tile
void hsp2g_plan_tasks(LandLaunch &L, int sub_pref, int resident_blocks, int warps_per_block, int *grid_out);
