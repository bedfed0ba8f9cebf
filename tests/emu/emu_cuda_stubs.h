// emu_cuda_stubs.h -- TEST DOUBLE ONLY.  Host stand-ins for the few CUDA runtime calls and device intrinsics that
// hspsquared_b200/csrc uses, so that the CPU test suite can compile the very same kernel and C-ABI sources with g++
// (-DHSP2G_HOST_EMU) and exercise host logic + kernel control flow without a GPU.  "Device" memory is malloc'd host
// memory, a "launch" is a loop over blockIdx/threadIdx, streams and events are no-ops.  libm here is glibc, so this
// double is bit-identical to the CPU oracle.  Never built or loaded by product code (see tests/emu/README.md).
#pragma once
#include <cmath>
#include <cstdlib>
#include <cstring>

#define __device__
#define __host__
#define __global__
#define __constant__
#define __shared__
#define __forceinline__ inline
#define __noinline__ inline
#define __grid_constant__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorEmu = 1 };
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0, cudaHostAllocPortable = 1,
       cudaHostAllocWriteCombined = 4, cudaHostRegisterDefault = 0, cudaHostRegisterPortable = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };
struct cudaDeviceProp {
    int major, minor;
};
struct EmuIdx {
    unsigned x;
};
extern thread_local EmuIdx threadIdx, blockIdx;
extern double ring[];  // the kernels' `extern __shared__ double ring[]`

inline const char *cudaGetErrorString(cudaError_t) { return "emulated CUDA error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { p->major = 10; p->minor = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorEmu; }
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaHostRegister(void *, size_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaHostUnregister(void *) { return cudaSuccess; }
inline cudaError_t cudaMemset(void *p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2D(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind) {
    for (size_t r = 0; r < h; ++r) memcpy((char *)d + r * dp, (const char *)s + r * sp, w);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind k,
                                     cudaStream_t) {
    return cudaMemcpy2D(d, dp, s, sp, w, h, k);
}
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) {
    memcpy(d, s, n);
    return cudaSuccess;
}
inline cudaError_t cudaMemset2D(void *p, size_t pitch, int v, size_t w, size_t h) {
    for (size_t r = 0; r < h; ++r) memset((char *)p + r * pitch, v, w);
    return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { return cudaMemset(p, v, n); }
inline cudaError_t cudaMemset2DAsync(void *p, size_t pitch, int v, size_t w, size_t h, cudaStream_t) {
    return cudaMemset2D(p, pitch, v, w, h);
}
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (void *)1; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (void *)1; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = (void *)1; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount };
enum { cudaErrorLaunchOutOfResources = 7 };
template <class K>
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, K, int, size_t) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) { *v = 1; return cudaSuccess; }
template <class K>
inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }

template <class T>
inline T __ldg(const T *p) { return *p; }
inline void __stcs(double *p, double v) { *p = v; }
inline void __stcs(float *p, float v) { *p = v; }
inline void __stcg(double *p, double v) { *p = v; }
template <class T>
inline T __ldcg(const T *p) { return *p; }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
    unsigned long long o = *p;
    *p = o + v;
    return o;
}
inline double __longlong_as_double(long long v) {
    double d;
    memcpy(&d, &v, sizeof d);
    return d;
}

#define HSP2G_LAUNCH(kernel, grid, block, smem, stream, arg)      \
    do {                                                          \
        for (unsigned b_ = 0; b_ < (unsigned)(grid); ++b_)        \
            for (unsigned t_ = 0; t_ < (unsigned)(block); ++t_) { \
                blockIdx.x = b_;                                  \
                threadIdx.x = t_;                                 \
                kernel(arg);                                      \
            }                                                     \
    } while (0)
