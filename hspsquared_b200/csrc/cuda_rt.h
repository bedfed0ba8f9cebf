// cuda_rt.h -- the one place the CUDA runtime header is pulled in.
// HSP2G_HOST_EMU is defined ONLY by tests/emu (a host-compiled test double of this library that lets the CPU test suite
// exercise the host logic and the kernels' control flow without a GPU).  The product build never defines it.
#pragma once
#ifdef HSP2G_HOST_EMU
#include "emu_cuda_stubs.h"
#else
#include <cuda_runtime.h>
#define HSP2G_LAUNCH(kernel, grid, block, smem, stream, arg) kernel<<<(grid), (block), (smem), (stream)>>>(arg)
#endif
