// common.h — internals shared by the translation units of libb200dsp.so
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>
#include <cstdio>
#include <atomic>
#include "b200dsp.h"

#define B200_API extern "C" __attribute__((visibility("default")))

struct B200Device {
    int ordinal = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 0;
    // rotating copy/compute streams + events for the *_host pipelined entry points
    static constexpr int kPipe = 3;
    cudaStream_t pipe[kPipe] = {nullptr, nullptr, nullptr};
    // scratch owned by the device object (grown on demand, reused by the drop-in host-pointer calls)
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    void *pinned = nullptr;
    size_t pinned_bytes = 0;
};

void b200_set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_b200_launches;
B200Device *b200_default_device();               // lazily opened, nullptr on failure
void *b200_scratch(B200Device *dev, size_t bytes);       // device scratch (>= bytes), nullptr on failure
void *b200_pinned(B200Device *dev, size_t bytes);        // pinned host scratch

#define B200_CUDA_OK(expr)                                                                      \
    do {                                                                                        \
        cudaError_t e_ = (expr);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            b200_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
            return B200_EEXTERNAL;                                                              \
        }                                                                                       \
    } while (0)

#define B200_LAUNCHED() (g_b200_launches.fetch_add(1, std::memory_order_relaxed))

static inline int b200_ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
