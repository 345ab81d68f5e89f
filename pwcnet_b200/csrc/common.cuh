// Shared helpers for libpwc_b200.so (sm_100a only).
#pragma once
#include <cstdlib>
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


// ---- programmatic dependent launch (PDL).  A kernel launched through launch_pdl may start while its predecessor in the
// stream is still running; it must execute pdl_wait() before it touches anything a predecessor writes (activations --
// weights and biases are only ever written by kernels that do not trigger early) and calls pdl_trigger() once to let ITS
// successor start.  Without the launch attribute (PWC_PDL=0) both are no-ops and the stream serialises as usual.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
inline int pdl_enabled() {
    static const int on = []() { const char* e = getenv("PWC_PDL"); return e ? atoi(e) : 1; }();
    return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// same, as thread-block clusters of `cluster` consecutive CTAs along x (grid.x must be a multiple of it)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, unsigned cluster,
                                      Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace pwc
