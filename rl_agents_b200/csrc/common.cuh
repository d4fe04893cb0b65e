// Shared host/device helpers of libb2planner (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/b2_planner.h"

namespace b2 {

void set_error(const char* fmt, ...);

#define B2_CUDA_CHECK(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            b2::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return B2_ERR_CUDA;                                                          \
        }                                                                                \
    } while (0)

#define B2_REQUIRE(cond, msg)                            \
    do {                                                 \
        if (!(cond)) {                                   \
            b2::set_error("invalid argument: %s", msg);  \
            return B2_ERR_INVALID;                       \
        }                                                \
    } while (0)

int sm_count();

__device__ __forceinline__ double warp_max_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w > v ? w : v;
    }
    return v;
}

}  // namespace b2
