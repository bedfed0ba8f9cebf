// emu_globals.cpp -- TEST DOUBLE: storage for the emulated launch indices and "shared memory" ring.
#include "emu_cuda_stubs.h"
thread_local EmuIdx threadIdx, blockIdx;
double ring[64 * 128 * 8];
