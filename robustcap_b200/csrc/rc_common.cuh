// Host-side plumbing shared by the .cu files: error reporting, launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <atomic>

#include "../../include/robustcap_b200.h"

void rc_set_error(const char* fmt, ...);
extern std::atomic<long long> g_rc_launches;

#define RC_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            rc_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return RC_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

#define RC_ARG(cond)                                                           \
    do {                                                                       \
        if (!(cond)) {                                                         \
            rc_set_error("%s:%d: bad argument: %s", __FILE__, __LINE__, #cond); \
            return RC_ERR_ARG;                                                 \
        }                                                                      \
    } while (0)

// every kernel launch of this library goes through RC_LAUNCH so bench.py can report how many ran
#define RC_LAUNCH(kernel, grid, block, smem, stream, ...)                       \
    do {                                                                        \
        kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__); \
        g_rc_launches.fetch_add(1, std::memory_order_relaxed);                  \
    } while (0)

#define RC_CHECK_LAUNCH() RC_CUDA(cudaGetLastError())

static inline int rc_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
