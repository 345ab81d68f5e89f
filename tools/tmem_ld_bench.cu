// tcgen05.ld (TMEM -> registers) throughput microbenchmark (sm_100a): one CTA per SM, W warps loop over
// 32x32b.xN loads of their own lane quadrant (warp % 4).  Prints clk per load and bytes/clk per SM.
// Also: packed fma.rn.f32x2 vs scalar fma.rn.f32 issue rate, and half-sector vs full-sector global store rate.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int N> struct Ld;
template <> struct Ld<16> {
    static __device__ __forceinline__ uint32_t go(uint32_t t) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(t));
        uint32_t x = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) x ^= r[i];
        return x;
    }
};
template <> struct Ld<32> {
    static __device__ __forceinline__ uint32_t go(uint32_t t) {
        uint32_t r[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                     "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                       "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                       "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(t));
        uint32_t x = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) x ^= r[i];
        return x;
    }
};
template <> struct Ld<64> {
    static __device__ __forceinline__ uint32_t go(uint32_t t) {
        uint32_t r[64];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                     "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
                     "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,"
                     "%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                       "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                       "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]),
                       "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]),
                       "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]),
                       "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
                     : "r"(t));
        uint32_t x = 0;
#pragma unroll
        for (int i = 0; i < 64; ++i) x ^= r[i];
        return x;
    }
};

template <int N>
__global__ void __launch_bounds__(512, 1) ld_bench(int iters, int depth, unsigned long long* out, uint32_t* sink) {
    __shared__ uint32_t slot;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    const int warp = threadIdx.x >> 5;
    const uint32_t tq = tm + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const unsigned long long t0 = clock64();
    // depth = loads in flight before a wait::ld
    for (int i = 0; i < iters; i += depth) {
        for (int d = 0; d < depth; ++d) acc ^= Ld<N>::go(tq + (((i + d) * N) & 511));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    const unsigned long long t1 = clock64();
    if (acc == 0x12345678u) sink[threadIdx.x] = acc;
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) out[warp] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

// ---- FFMA vs FFMA2 issue rate: 8 warps per SMSP-equivalent, long dependent-free chains
__global__ void __launch_bounds__(256, 1) ffma_bench(int iters, int packed, unsigned long long* out, float* sink) {
    float a[16], b = 1.0001f + threadIdx.x * 1e-7f, c = 0.5f;
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = i * 0.25f + threadIdx.x;
    __syncthreads();
    const unsigned long long t0 = clock64();
    if (packed) {
        unsigned long long bb, cc;
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
        unsigned long long p[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(bb), "l"(cc));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(p[i]));
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
        }
    }
    const unsigned long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 1.2345f) sink[threadIdx.x] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
}

// ---- global store patterns: each lane writes 324-byte runs at a 592-byte pixel pitch (the concat slot of level 2)
//      mode 0: lane = pixel, 20 x STG.128 + 1 x STG.32 (half-sector writes);  mode 1: coalesced (lane = float4 unit of a run)
__global__ void __launch_bounds__(256) st_bench(float* out, size_t n_pix, int mode) {
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (size_t p0 = warp * 32; p0 < n_pix; p0 += nw * 32) {
        if (mode == 0) {
            float* q = out + (p0 + lane) * 148;
#pragma unroll
            for (int k = 0; k < 20; ++k) *reinterpret_cast<float4*>(q + 4 * k) = make_float4(1.f, 2.f, 3.f, (float)k);
            q[80] = 5.f;
        } else {
            int pix = lane / 21, k = lane - pix * 21;
#pragma unroll 3
            for (int m = 0; m < 21; ++m) {
                float* q = out + (p0 + pix) * 148 + 4 * k;
                if (k < 20) *reinterpret_cast<float4*>(q) = make_float4(1.f, 2.f, 3.f, (float)k); else *q = 5.f;
                k += 11; pix += 1;
                if (k >= 21) { k -= 21; pix += 1; }
            }
        }
    }
}

int main() {
    unsigned long long* d; cudaMalloc(&d, 16 * 8);
    uint32_t* sink; cudaMalloc(&sink, 4096);
    const int iters = 2048;
    printf("tcgen05.ld 32x32b.xN, %d loads per warp, 148 CTAs; W warps (warp %% 4 = lane quadrant)\n", iters);
    for (int depth : {1, 2, 4})
        for (int W : {1, 4, 8, 16}) {
            for (int N : {16, 32, 64}) {
                for (int rep = 0; rep < 2; ++rep) {
                    if (N == 16) ld_bench<16><<<148, W * 32>>>(iters, depth, d, sink);
                    if (N == 32) ld_bench<32><<<148, W * 32>>>(iters, depth, d, sink);
                    if (N == 64) ld_bench<64><<<148, W * 32>>>(iters, depth, d, sink);
                }
                unsigned long long h[16]; cudaMemcpy(h, d, 16 * 8, cudaMemcpyDeviceToHost);
                cudaError_t e = cudaGetLastError();
                unsigned long long mx = 0; for (int w = 0; w < W; ++w) mx = h[w] > mx ? h[w] : mx;
                printf("depth %d  warps %2d  x%-3d : %7.1f clk/load/warp   %7.1f B/clk/SM %s\n", depth, W, N, (double)mx / iters,
                       (double)W * iters * N * 128.0 / mx, e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
        }
    float* fs; cudaMalloc(&fs, 4096);
    for (int packed = 0; packed < 2; ++packed) {
        ffma_bench<<<148, 256>>>(4096, packed, d, fs);
        ffma_bench<<<148, 256>>>(4096, packed, d, fs);
        unsigned long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        // 8 warps x 16 FMA-lanes-worth per iteration per thread
        printf("%s: %.2f clk per iteration (16 fp32 FMAs per thread, 8 warps/SM) -> %.1f FMA/clk/SM\n", packed ? "fma.rn.f32x2" : "fma.rn.f32  ",
               (double)h / 4096, 256.0 * 16 * 4096 / h);
    }
    const size_t n_pix = (size_t)8 * 112 * 256;
    float* o; cudaMalloc(&o, n_pix * 148 * 4);
    char* fl; cudaMalloc(&fl, 256 << 20);
    for (int mode = 0; mode < 2; ++mode) {
        float best = 1e9;
        for (int rep = 0; rep < 5; ++rep) {
            cudaMemset(fl, rep, 256 << 20);
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            cudaEventRecord(a); st_bench<<<148 * 8, 256>>>(o, n_pix, mode); cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best;
        }
        printf("store 74 MB of 324-byte runs at pitch 592 B, %s: %.1f us (%.0f GB/s)\n", mode ? "coalesced float4 units" : "lane = pixel (half-sector STG.128)",
               best * 1e3, n_pix * 324.0 / best / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
