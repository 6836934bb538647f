// Shared helpers of libagx.so (error reporting, launch checks, small device utilities).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/agx.h"

namespace agx {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch();     // bumps the process-wide kernel launch counter (agx_launch_count)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs

// Side streams for INDEPENDENT launches of one entry point (the latency-bound few-CTA GEMM classes
// next to the tensor-core launch).  fork(): the side stream waits for everything queued on `main`
// so far; join(): `main` waits for the side stream.  Works under stream capture (the branch becomes
// a parallel path of the captured graph).  One pair per device, created on first use; env
// AGX_NO_SIDE_STREAMS=1 makes fork() hand back `main` itself (serial launches, A/B switch).
struct SideStream {
    cudaStream_t main, side;
    bool forked;
    int which_;
    int fork(cudaStream_t main_stream, int which = 0);
    int join();
};

}  // namespace agx

#define AGX_CHECK_ARG(cond, ...)                       \
    do {                                               \
        if (!(cond)) {                                 \
            agx::set_error(__VA_ARGS__);               \
            return AGX_ERR_INVALID;                    \
        }                                              \
    } while (0)

#define AGX_CUDA(call)                                             \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return agx::cuda_fail(e_, #call);   \
    } while (0)

#define AGX_LAUNCH_CHECK(name)                                      \
    do {                                                            \
        cudaError_t e_ = cudaGetLastError();                        \
        if (e_ != cudaSuccess) return agx::cuda_fail(e_, name);     \
        agx::count_launch();                                        \
    } while (0)

// 128-bit read-only global load that does not allocate in L1 (streaming gathers)
__device__ __forceinline__ float4 ldg_f4(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
