// Shared helpers for libpwc_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "../../include/pwc_b200.h"

namespace pwc {

void set_error(const char* fmt, ...);
int sm_count();   // SMs of the current device (abi.cu)

__host__ __device__ inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Returns from the enclosing extern "C" function with the launch error, if any.
#define PWC_CHECK_LAUNCH(name)                                                    \
    do {                                                                          \
        cudaError_t e__ = cudaGetLastError();                                     \
        if (e__ != cudaSuccess) {                                                 \
            pwc::set_error("%s: %s", name, cudaGetErrorString(e__));              \
            return (int)e__;                                                      \
        }                                                                         \
    } while (0)

#define PWC_REQUIRE(cond, code, ...)                                              \
    do {                                                                          \
        if (!(cond)) {                                                            \
            pwc::set_error(__VA_ARGS__);                                          \
            return (code);                                                        \
        }                                                                         \
    } while (0)

__device__ __forceinline__ float leaky(float v, float alpha) { return fmaxf(alpha * v, v); }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

}  // namespace pwc
