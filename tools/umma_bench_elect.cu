// tcgen05.mma issue/throughput microbenchmark (sm_100a): one CTA per SM, operands = zeros in shared memory.
//   ./umma_bench        prints clk per MMA for N in {16..256}, operand layout {128B rows (32 B used per K step), 64B rows},
//   accumulator {same, rotating over 2/4}.  M = 128, kind::f16, K = 16, cta_group::1.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__global__ void __launch_bounds__(128, 1) bench(int N, int layout, int nacc, int iters, unsigned long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    uint32_t is_leader = 0;
    if (threadIdx.x < 32) asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(is_leader));
    if (is_leader) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        // layout 0: rows of 128 B, 128B swizzle (SBO 1024); layout 1: rows of 64 B, 64B swizzle (SBO 512)
        const uint64_t hi = layout != 1 ? (((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61))
                                        : (((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61));
        const uint64_t ad = hi | (uint64_t)(((base >> 4) & 0x3FFF) | (1u << 16));
        const uint64_t bd = hi | (uint64_t)((((base + 16384) >> 4) & 0x3FFF) | (1u << 16));
        const unsigned long long t0 = clock64();
        const uint32_t mask = nacc - 1;
        if (layout == 2) {   // A start address shifted by whole 128-byte rows per MMA (the halo kernel's tap shifts: 0,1,2,130,131,...)
#pragma unroll 8
            for (int i = 0; i < iters; ++i) mma(tm + (i & mask) * N, ad + 8 * ((i & 7) * 17), bd + 2 * (i & 1), idesc, 1u);
        } else {
#pragma unroll 8
        for (int i = 0; i < iters; ++i) mma(tm + (i & mask) * N, ad + 2 * (i & 1), bd + 2 * (i & 1), idesc, 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        const unsigned long long t1 = clock64();
        asm volatile("{\n\t.reg .pred P1;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
        const unsigned long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}
int main() {
    unsigned long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int iters = 512;
    printf("M=128 K=16 kind::f16, %d MMAs issued by one thread, 148 CTAs; clk per MMA (issue loop / until complete)\n", iters);
    for (int layout = 0; layout < 3; ++layout)
        for (int nacc = 1; nacc <= 4; nacc *= 4)
            for (int N : {16, 32, 64, 128, 256}) {
                if (nacc * N > 512) continue;
                bench<<<148, 128, 64 * 1024>>>(N, layout, nacc, iters, d);
                bench<<<148, 128, 64 * 1024>>>(N, layout, nacc, iters, d);
                unsigned long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                cudaError_t e = cudaGetLastError();
                printf("layout %s  accumulators %d  N %3d : issue %6.1f  complete %6.1f  (floor %d) %s\n", layout == 1 ? "64B rows " : (layout == 2 ? "128B shift" : "128B rows"), nacc, N,
                       (double)h[0] / iters, (double)h[1] / iters, 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    return 0;
}
