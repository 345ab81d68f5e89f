// SM -> HBM store-path microbenchmark (sm_100a): 148 persistent CTAs write a 1 GiB buffer (>> L2) in 4 KB chunks
//   mode 0: STG.128 from registers, W warps per CTA          mode 1: cp.async.bulk shared -> global, one thread, 4 KB copies
//   mode 2: STG.128 of 352-byte runs at a 608-byte pitch (the concat-slot head the cost-volume kernel writes)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(1024, 1) st_kernel(float4* out, size_t n_chunks, int mode, int pitch) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, nth = blockDim.x;
    if (mode == 0) {
        // chunk = 4 KB = 256 float4; thread t writes float4 t, t + nth, ...
        for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
            float4* p = out + c * 256;
            for (int i = tid; i < 256; i += nth) p[i] = make_float4(1.f, 2.f, 3.f, (float)i);
        }
    } else if (mode == 1) {
        for (int i = tid; i < 4096 / 4; i += nth) reinterpret_cast<float*>(smem)[i] = (float)i;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            int inflight = 0;
            for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + c * 256), "r"(smem_u32(smem)), "r"(4096) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (++inflight >= 8) { asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); inflight = 4; }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else if (mode == 3) {
        // 352-byte bulk copies at a 608-byte pitch: one elected thread, one copy per pixel
        for (int i = tid; i < 4096 / 4; i += nth) reinterpret_cast<float*>(smem)[i] = (float)i;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            const size_t n_pix = n_chunks * 4096 / 608;
            int inflight = 0;
            for (size_t p0 = (size_t)blockIdx.x * 8; p0 + 8 <= n_pix; p0 += (size_t)gridDim.x * 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<char*>(out) + (p0 + j) * 608),
                                 "r"(smem_u32(smem) + j * 352), "r"(352) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (++inflight >= 8) { asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); inflight = 4; }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else {
        // 32 pixels per warp iteration: 22 float4 units each (352 B) at `pitch` bytes; unit u = lane + 32 m
        const int warp = tid >> 5, lane = tid & 31, nw = nth >> 5;
        const size_t n_pix = n_chunks * 4096 / pitch;
        for (size_t p0 = ((size_t)blockIdx.x * nw + warp) * 32; p0 + 32 <= n_pix; p0 += (size_t)gridDim.x * nw * 32) {
            char* base = reinterpret_cast<char*>(out) + p0 * pitch;
#pragma unroll
            for (int m = 0; m < 22; ++m) {
                const int u = lane + 32 * m, pix = u / 22, k = u - pix * 22;
                *reinterpret_cast<float4*>(base + (size_t)pix * pitch + 16 * k) = make_float4(1.f, 2.f, 3.f, (float)m);
            }
        }
    }
}

int main() {
    const size_t bytes = 1ull << 30;
    float4* buf; cudaMalloc(&buf, bytes);
    cudaFuncSetAttribute(st_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto run = [&](int mode, int threads, const char* name, double useful_frac, int pitch = 608) {
        float best = 1e9;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(a);
            st_kernel<<<148, threads, 8192>>>(buf, bytes / 4096, mode, pitch);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); if (rep && ms < best) best = ms;
        }
        const double gb = bytes * useful_frac / 1e9;
        printf("%-62s threads %4d: %7.3f ms  %7.1f GB/s  %5.1f B/clk/SM @1.9GHz  %s\n", name, threads, best, gb / (best * 1e-3),
               bytes * useful_frac / (best * 1e-3) / 148 / 1.9e9, cudaGetErrorString(cudaGetLastError()));
    };
    for (int th : {64, 128, 256, 512, 1024}) run(0, th, "STG.128 contiguous 4 KB chunks", 1.0);
    run(1, 32, "cp.async.bulk smem->global 4 KB, one thread, 8 in flight", 1.0);
    for (int th : {64, 128, 256, 512}) run(2, th, "STG.128 352-byte runs at 608-byte pitch", 352.0 / 608.0);
    run(3, 32, "cp.async.bulk 352 B per pixel at 608-byte pitch, one thread", 352.0 / 608.0);
    for (int th : {64, 256}) run(2, th, "STG.128 352-byte runs at 640-byte pitch", 352.0 / 640.0, 640);
    for (int th : {64, 256}) run(2, th, "STG.128 352-byte runs at 384-byte pitch (dense 96-word rows)", 352.0 / 384.0, 384);
    for (int th : {64, 256}) run(2, th, "STG.128 352-byte runs at 352-byte pitch (dense 88-word rows)", 1.0, 352);
    // reference: cudaMemsetAsync
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a); cudaMemsetAsync(buf, rep, bytes); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (rep && ms < best) best = ms;
    }
    printf("cudaMemsetAsync 1 GiB: %.3f ms  %.1f GB/s\n", best, bytes / 1e9 / (best * 1e-3));
    return 0;
}
