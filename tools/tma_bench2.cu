// Microbenchmark 2: conv-like TMA pattern.  Per stage: one 4-D box {32 ch,16 px,8 rows,1} of an NHWC fp32 tensor
// (B=8,H=112,W=256,C=128) at a tap offset + two 3-D weight boxes {32 fp16(64B),128 rows,1} -> 32 KB, S stages in flight.
// Reports per-stage period and issue->complete latency, optionally with 4 warps hammering shared memory (LDS/STS)
// to emulate the converter / UMMA operand traffic.
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(192, 1) k(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, int iters, int stages, int hammer, int wonly, unsigned long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[8];
    __shared__ int stop;
    const uint32_t base = (s32(smem) + 1023u) & ~1023u;
    if (threadIdx.x == 0) {
        stop = 0;
        for (int i = 0; i < stages; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int tile = blockIdx.x;   // 16x8 tiles: 16 x 14 per image
        unsigned long long t0 = clock64(), lat = 0, issue_t[8];
        for (int it = 0; it < iters + stages; ++it) {
            const int s = it % stages;
            if (it >= stages) {
                const uint32_t ph = ((it / stages) - 1) & 1;
                asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(s32(&full[s])), "r"(ph) : "memory");
                lat += clock64() - issue_t[s];
            }
            if (it < iters) {
                const int tt = (tile + (it / 36) * 148) % (16 * 14 * 8);
                const int tx = tt % 16, ty = (tt / 16) % 14, b = tt / 224;
                const int st = it % 36, tap = st / 4, kc = st % 4, ky = tap / 3, kx = tap % 3;
                const uint32_t dst = base + s * 32768;
                issue_t[s] = clock64();
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(wonly ? 16384 : 32768) : "memory");
                if (!wonly)
                    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                                 ::"r"(dst), "l"(&tmX), "r"(s32(&full[s])), "r"(kc * 32), "r"(tx * 16 + kx - 1), "r"(ty * 8 + ky - 1), "r"(b) : "memory");
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                             ::"r"(dst + 16384), "l"(&tmW), "r"(s32(&full[s])), "r"(kc * 32), "r"(0), "r"(tap) : "memory");
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                             ::"r"(dst + 24576), "l"(&tmW), "r"(s32(&full[s])), "r"(kc * 32), "r"(0), "r"(9 + tap) : "memory");
            }
        }
        out[blockIdx.x * 2] = clock64() - t0;
        out[blockIdx.x * 2 + 1] = lat;
        stop = 1;
    } else if (threadIdx.x >= 64 && hammer) {
        // shared-memory traffic generator on a private region (beyond the stage buffers)
        float4* p = reinterpret_cast<float4*>(smem + 1024 + 6 * 32768);
        float4 acc = make_float4(0, 0, 0, 0);
        const int t = threadIdx.x - 64;
        while (!*(volatile int*)&stop) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(s32(p + t + 128 * i)));
                acc.x += v.x; acc.y += v.y;
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(s32(p + t + 128 * i)), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w) : "memory");
            }
        }
        if (acc.x == 12345.f) out[0] = 1;
    }
}
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    void* fp; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    Enc enc = (Enc)fp;
    const int B = 8, H = 112, W = 256, C = 128;
    float* x; cudaMalloc(&x, (size_t)B * H * W * C * 4); cudaMemset(x, 0, (size_t)B * H * W * C * 4);
    void* w; cudaMalloc(&w, 18 * 128 * 128 * 2); cudaMemset(w, 0, 18 * 128 * 128 * 2);
    unsigned long long* out; cudaMalloc(&out, 148 * 16);
    CUtensorMap tmX, tmW;
    { cuuint64_t dims[4] = {C, W, H, B}; cuuint64_t st[3] = {C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4}; cuuint32_t box[4] = {32, 16, 8, 1}; cuuint32_t es[4] = {1, 1, 1, 1};
      enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
    { cuuint64_t dims[3] = {128, 128, 18}; cuuint64_t st[2] = {256, 128 * 256}; cuuint32_t box[3] = {32, 128, 1}; cuuint32_t es[3] = {1, 1, 1};
      enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, w, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
    const size_t smem = 1024 + 6 * 32768 + 16384;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int iters = 36 * 12;
    for (int wonly = 0; wonly < 2; ++wonly) for (int hammer = 0; hammer < 2; ++hammer) for (int stages = 2; stages <= 6; stages += 2) {
        for (int rep = 0; rep < 2; ++rep) k<<<148, 192, smem>>>(tmX, tmW, iters, stages, hammer, wonly, out);
        cudaError_t err = cudaGetLastError(); if (err == cudaSuccess) err = cudaDeviceSynchronize();
        unsigned long long h[296]; cudaMemcpy(h, out, 296 * 8, cudaMemcpyDeviceToHost);
        double per = 0, lat = 0; for (int i = 0; i < 148; ++i) { per += h[2 * i]; lat += h[2 * i + 1]; } per /= 148.0 * iters; lat /= 148.0 * iters;
        printf("%s stages %d hammer %d: period %.0f clk/stage (%.1f B/clk/SM), issue->complete latency %.0f clk (%s)\n", wonly ? "weights only (16KB)" : "A 4-D + 2 W (32KB)  ", stages, hammer, per, (wonly ? 16384.0 : 32768.0) / per, lat, cudaGetErrorString(err));
    }
    return 0;
}
