// Microbenchmark: per-SM TMA fill throughput into shared memory vs request shape (B200).
//   mode 0: 2-D tensor box {32 fp32 (128 B), 128 rows}, rows 512 B apart in global (like an NHWC channel slice), SWIZZLE_128B
//   mode 1: 2-D tensor box {16 fp32 (64 B), 256 rows}, SWIZZLE_64B            (same 16 KB per load)
//   mode 2: 1-D bulk copy of 16 KB contiguous
//   mode 3: 2-D tensor box {64 fp32 (256 B), 64 rows}, no swizzle
//   mode 4: 2-D box {32 fp32, 128 rows} rows CONTIGUOUS in global (128 B apart)
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap tm, const float* src, int mode, int iters, int stages, unsigned long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[8];
    const uint32_t base = (s32(smem) + 1023u) & ~1023u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t0 = clock64();
        for (int it = 0; it < iters + stages; ++it) {
            const int s = it % stages;
            if (it >= stages) {   // wait for the load issued `stages` iterations ago into this slot
                const uint32_t ph = ((it / stages) - 1) & 1;
                asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(s32(&full[s])), "r"(ph) : "memory");
            }
            if (it < iters) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(16384) : "memory");
                const uint32_t dst = base + s * 16384;
                const int row0 = ((blockIdx.x * 37 + it) * 256) % 65536;   // walk over a 128 MB region (L2-resident after warm-up)
                if (mode == 2) {
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(dst), "l"(src + (size_t)row0 * 128), "r"(16384), "r"(s32(&full[s])) : "memory");
                } else {
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                 ::"r"(dst), "l"(&tm), "r"(s32(&full[s])), "r"(0), "r"(row0) : "memory");
                }
            }
        }
        out[blockIdx.x] = clock64() - t0;
    }
}
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    void* fp; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    Enc enc = (Enc)fp;
    float* src; cudaMalloc(&src, (size_t)70000 * 128 * 4 + (1 << 20)); cudaMemset(src, 0, (size_t)70000 * 128 * 4);
    unsigned long long* out; cudaMalloc(&out, 148 * 8);
    const int iters = 2000;
    const char* names[5] = {"2D box 128B rows x128 (stride 512B, SW128)", "2D box 64B rows x256 (stride 512B, SW64)", "1D bulk 16KB contiguous", "2D box 256B rows x64 (stride 512B, no swizzle)", "2D box 128B rows x128 (contiguous rows, SW128)"};
    for (int mode = 0; mode < 5; ++mode) for (int stages = 4; stages <= 8; stages += 4) {
        CUtensorMap tm; 
        cuuint64_t dims[2] = {128, 66000}; cuuint64_t strides[1] = {512}; cuuint32_t es[2] = {1, 1};
        cuuint32_t box[2] = {32, 128}; CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
        if (mode == 1) { box[0] = 16; box[1] = 256; sw = CU_TENSOR_MAP_SWIZZLE_64B; }
        if (mode == 3) { box[0] = 64; box[1] = 64; sw = CU_TENSOR_MAP_SWIZZLE_NONE; }
        if (mode == 4) { dims[0] = 32; dims[1] = 66000 * 4; strides[0] = 128; }
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed mode %d: %d\n", mode, (int)r); continue; }
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384 + 1024);
        for (int rep = 0; rep < 2; ++rep) k<<<148, 128, 8 * 16384 + 1024>>>(tm, src, mode, iters, stages, out);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0); k<<<148, 128, 8 * 16384 + 1024>>>(tm, src, mode, iters, stages, out); cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        unsigned long long h[148]; cudaMemcpy(h, out, 148 * 8, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        printf("%-50s stages %d: %.1f clk per 16KB load per SM = %.1f B/clk/SM ; chip %.2f TB/s (%s)\n", names[mode], stages, avg / iters, 16384.0 * iters / avg, 148.0 * iters * 16384 / (ms * 1e-3) / 1e12, cudaGetErrorString(err));
    }
    return 0;
}
