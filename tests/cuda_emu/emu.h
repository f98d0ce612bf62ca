// emu.h — TEST INFRASTRUCTURE ONLY.  A few dozen lines that let the device code of the kernels written without GPU access
// compile and run on the host, one CUDA thread after the other (a real OS thread per lane where a kernel uses a warp
// collective), so their arithmetic and indexing can be compared with the checker before they ever see a GPU.
// It emulates only what those kernels use: built-in index variables, __ldg, the round-to-nearest float intrinsics, min / max,
// uint4, __reduce_xor_sync.  It does not model memory spaces, races or performance.
#pragma once
#include <cstdint>
#include <cstddef>
#include <algorithm>
#include <barrier>
#include <thread>
#include <vector>
#include <functional>

struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct uint4 { unsigned x, y, z, w; };
struct float2 { float x, y; };
static inline float2 make_float2(float a, float b) { float2 r = { a, b }; return r; }
static thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __constant__ static const
#define __restrict__

template <typename T> static inline T __ldg(const T *p) { return *p; }
using std::min;
using std::max;
static inline float  __fmul_rn(float a, float b)   { volatile float r = a * b; return r; }      // volatile: no contraction into an fma
static inline float  __fadd_rn(float a, float b)   { volatile float r = a + b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }

// warp collective for kernels launched with emu_launch_warps: every lane is an OS thread, the barrier makes them meet
struct EmuWarp { std::barrier<> *bar; unsigned slots[32]; };
static thread_local EmuWarp *emu_warp = nullptr;
static inline unsigned __reduce_xor_sync(unsigned, unsigned v)
{
    EmuWarp *w = emu_warp;
    w->slots[threadIdx.x & 31] = v;
    w->bar->arrive_and_wait();
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r ^= w->slots[i];
    w->bar->arrive_and_wait();
    return r;
}

// sequential launch: kernels without collectives or shared memory
template <typename F> static void emu_launch(dim3 grid, dim3 block, F body)
{
    gridDim = grid; blockDim = block;
    for (unsigned bz = 0; bz < grid.z; bz++) for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++) {
        blockIdx = dim3(bx, by, bz);
        for (unsigned tz = 0; tz < block.z; tz++) for (unsigned ty = 0; ty < block.y; ty++) for (unsigned tx = 0; tx < block.x; tx++) {
            threadIdx = dim3(tx, ty, tz);
            body();
        }
    }
}

// one OS thread per lane, warp by warp (1-D blocks): for kernels that use __reduce_xor_sync
template <typename F> static void emu_launch_warps(dim3 grid, dim3 block, F body)
{
    for (unsigned bx = 0; bx < grid.x; bx++)
        for (unsigned w0 = 0; w0 < block.x; w0 += 32) {
            std::barrier<> bar(32);
            EmuWarp warp{ &bar, { 0 } };
            std::vector<std::thread> lanes;
            for (unsigned l = 0; l < 32; l++)
                lanes.emplace_back([&, l] {
                    gridDim = grid; blockDim = block; blockIdx = dim3(bx, 0, 0); threadIdx = dim3(w0 + l, 0, 0);
                    emu_warp = &warp;
                    body();
                });
            for (auto &t : lanes) t.join();
        }
}
