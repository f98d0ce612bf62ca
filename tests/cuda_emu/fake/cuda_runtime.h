// cuda_runtime.h — TEST INFRASTRUCTURE ONLY: a stand-in for the CUDA runtime so that the HOST side of the .cu files that contain
// no inline PTX (entry points, argument checks, chunk loops, copies, launch geometry) can be compiled with g++ and run on the
// CPU test tier.  "Device memory" is host memory; every call completes synchronously; kernels run under emu.h.
#pragma once
#include "../emu.h"
#include <cstdlib>
#include <cstring>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1 };
typedef struct FakeStream *cudaStream_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };

static inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "fake error"; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
template <typename T> static inline cudaError_t cudaMalloc(T **p, size_t n) { *p = (T *)malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorInvalidValue; }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t)
{
    for (size_t r = 0; r < h; r++) memcpy((char *)d + r * dp, (const char *)s + r * sp, w);
    return cudaSuccess;
}

// what `kernel<<<grid, block, smem, stream>>>(args)` is rewritten into by tests/test_cuda_emu.py
// mode 0: threads one after the other; 1: an OS thread per lane (warp collectives); 2: an OS thread per thread of a block (__syncthreads)
template <typename F> static void emu_cfg_launch(int mode, F body, dim3 grid, dim3 block, size_t smem = 0, cudaStream_t = nullptr)
{
    if (mode == 1) emu_launch_warps(grid, block, body);
    else if (mode == 2) emu_launch_blocks(grid, block, smem, body);
    else emu_launch(grid, block, body);
}
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) { *v = 2; return cudaSuccess; }
template <typename K> static inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
