// tests/emu/fake/cuda_runtime.h -- TEST INFRASTRUCTURE ONLY.
// A single-threaded host stand-in for the handful of CUDA runtime calls and device intrinsics the
// library uses, so that the *same kernel sources* can be compiled with g++ and stepped through on the
// GPU-less authoring box (tests/emu/build_emu.py).  It is never built into, loaded by or reachable
// from compton2d_b200/: the product has no CPU path.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#define C2D_EMU 1
#define __device__
#define __global__
#define __grid_constant__
#define __host__
#define __constant__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __restrict__

struct EmuDim { unsigned x = 0, y = 0, z = 0; };
extern EmuDim blockIdx, threadIdx, blockDim, gridDim;

typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1 };
struct cudaDeviceProp { int major = 10, minor = 0, multiProcessorCount = 148; };

inline const char *cudaGetErrorString(cudaError_t) { return "emu"; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return 0; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { *p = cudaDeviceProp(); return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, int) { *s = nullptr; return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
template <class T> inline cudaError_t cudaMalloc(T **p, size_t n) { *p = (T *)calloc(n ? n : 1, 1); return 0; }
inline cudaError_t cudaFree(void *p) { free(p); return 0; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t)
{
    for (size_t i = 0; i < h; i++) memcpy((char *)d + i * dp, (const char *)s + i * sp, w);
    return 0;
}
inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return 0; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = nullptr; return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0; return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaFuncSetAttribute(const void *, int, int) { return 0; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

template <class T> inline T atomicAdd(T *a, T v) { T o = *a; *a = o + v; return o; }
inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline void __syncthreads() {}
inline void __syncwarp(unsigned = 0xffffffffu) {}
