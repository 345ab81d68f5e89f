// Warp-level mma.sync / ldmatrix throughput on sm_100a (the legacy tensor-core path next to tcgen05).
// Decides whether a register-fragment cost volume (band of 16 px x 24 candidates per m16n8k16, diagonal extraction
// from the D fragments, no TMEM read-back) can beat the 57 us CUDA-core kernel:  level 2 needs 4.75 GMAC of
// m16n8k16 f16 (3 x fp16 split, 37.5 % of each D tile useful), i.e. >= 1100 MAC/clk/SM for 15 us.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_sync_bench.bin tools/mma_sync_bench.cu
//   ./tools/mma_sync_bench.bin   ->  MAC/clk/SM for {f16 m16n8k16, tf32 m16n8k8} x {4, 8, 16 warps/SM} x ILP {1, 2, 4, 8},
//                                    and with one ldmatrix.x4 per {3, 6, 12} MMAs (shared-memory operand traffic).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ void mma_f16(float* d, const uint32_t* a, const uint32_t* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_tf32(float* d, const uint32_t* a, const uint32_t* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t* r, uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}

// kind 0: f16 m16n8k16 (2048 MAC), 1: tf32 m16n8k8 (1024 MAC).  ILP independent accumulators per warp.
// ld_every > 0: one ldmatrix.x4 (512 B of shared memory) refreshes the B fragments every `ld_every` MMAs.
template <int KIND, int ILP>
__global__ void bench(int iters, int ld_every, float* sink, unsigned long long* clk) {
    __shared__ __align__(16) uint32_t tile[32 * 40];
    for (int i = threadIdx.x; i < 32 * 40; i += blockDim.x) tile[i] = 0;
    __syncthreads();
    float d[ILP][4];
    uint32_t a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < ILP; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
    const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(tile + (threadIdx.x & 31) * 40 % (32 * 36));
    __syncthreads();
    const unsigned long long t0 = clock64();
    int since = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (ld_every > 0 && ++since >= ld_every) { ldmatrix_x4(b, saddr); since = 0; }
            if (KIND == 0) mma_f16(d[i], a, b); else mma_tf32(d[i], a, b);
        }
    }
    const unsigned long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
    if (s == 123.f) sink[0] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int KIND, int ILP>
static void run(int warps, int ld_every, float* sink, unsigned long long* clk, int n_sm) {
    const int iters = 4096;
    bench<KIND, ILP><<<n_sm, warps * 32>>>(64, ld_every, sink, clk);
    bench<KIND, ILP><<<n_sm, warps * 32>>>(iters, ld_every, sink, clk);
    cudaDeviceSynchronize();
    unsigned long long h[256];
    cudaMemcpy(h, clk, sizeof(unsigned long long) * n_sm, cudaMemcpyDeviceToHost);
    unsigned long long worst = 0;
    for (int i = 0; i < n_sm; ++i) worst = h[i] > worst ? h[i] : worst;
    const double mac = (KIND == 0 ? 2048.0 : 1024.0) * iters * ILP * warps;
    printf("%s warps/SM %2d ILP %d ldmatrix every %2d MMAs: %7.1f MAC/clk/SM  (%.2f clk per MMA per SM)\n", KIND == 0 ? "f16 m16n8k16" : "tf32 m16n8k8",
           warps, ILP, ld_every, mac / (double)worst, (double)worst / ((double)iters * ILP * warps));
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int n_sm = prop.multiProcessorCount;
    float* sink; unsigned long long* clk;
    cudaMalloc(&sink, 4); cudaMalloc(&clk, sizeof(unsigned long long) * 256);
    printf("%s, %d SMs\n", prop.name, n_sm);
    for (int warps : {4, 8, 16}) {
        run<0, 1>(warps, 0, sink, clk, n_sm); run<0, 2>(warps, 0, sink, clk, n_sm);
        run<0, 4>(warps, 0, sink, clk, n_sm); run<0, 8>(warps, 0, sink, clk, n_sm);
    }
    for (int warps : {8, 16}) { run<1, 4>(warps, 0, sink, clk, n_sm); run<1, 8>(warps, 0, sink, clk, n_sm); }
    for (int ld : {3, 6, 12}) { run<0, 8>(8, ld, sink, clk, n_sm); run<0, 8>(16, ld, sink, clk, n_sm); }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
