// Host-side plumbing shared by the .cu files: error reporting, launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <string.h>
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

// Programmatic dependent launch (PDL): the kernel may be scheduled while its predecessor in the stream drains, which hides the
// launch / scheduling latency of the ~11 kernel boundaries of a frame.  Every kernel launched this way starts with rc_pdl_wait()
// (griddepcontrol.wait: the predecessor has completed and flushed) before it touches global memory, and calls rc_pdl_trigger()
// (griddepcontrol.launch_dependents) so that its own successor may be scheduled early.  RC_NO_PDL=1 turns the attribute off.
#ifdef __CUDACC__
__device__ __forceinline__ void rc_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void rc_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool rc_pdl_enabled();

template <typename... KArgs, typename... Args>
static inline cudaError_t rc_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, void* stream, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = rc_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#define RC_LAUNCH_PDL(kernel, grid, block, smem, stream, ...)                                      \
    do {                                                                                           \
        cudaError_t e_ = rc_launch_pdl(kernel, dim3(grid), dim3(block), (smem), (stream), __VA_ARGS__); \
        if (e_ != cudaSuccess) {                                                                   \
            rc_set_error("%s:%d: launch of %s -> %s", __FILE__, __LINE__, #kernel, cudaGetErrorString(e_)); \
            return RC_ERR_CUDA;                                                                    \
        }                                                                                          \
        g_rc_launches.fetch_add(1, std::memory_order_relaxed);                                     \
    } while (0)
#endif

#define RC_CHECK_LAUNCH() RC_CUDA(cudaGetLastError())

static inline int rc_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
