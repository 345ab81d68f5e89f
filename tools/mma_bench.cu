// Microbenchmark: legacy mma.sync.m16n8k8 tf32 throughput and FFMA / FFMA2 throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void mma_tf32(float* out, int iters) {
    unsigned a[4] = {threadIdx.x, 2, 3, 4}, b[2] = {5, 6};
    float c[8][4] = {};
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0; for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void mma_bf16(float* out, int iters) {
    unsigned a[4] = {threadIdx.x, 2, 3, 4}, b[2] = {5, 6};
    float c[8][4] = {};
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0; for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void ffma1(float* out, int iters) {
    float c[32]; float a = threadIdx.x * 1e-3f, b = 1.0001f;
#pragma unroll
    for (int j = 0; j < 32; ++j) c[j] = j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 32; ++j) c[j] = fmaf(c[j], b, a);
    }
    float s = 0; for (int j = 0; j < 32; ++j) s += c[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void ffma2(float* out, int iters) {
    float2 c[16]; float2 a = make_float2(threadIdx.x * 1e-3f, 0.5f), b = make_float2(1.0001f, 0.9999f);
#pragma unroll
    for (int j = 0; j < 16; ++j) c[j] = make_float2(j, -j);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            unsigned long long cc = *reinterpret_cast<unsigned long long*>(&c[j]);
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(cc) : "l"(*reinterpret_cast<unsigned long long*>(&b)), "l"(*reinterpret_cast<unsigned long long*>(&a)));
            c[j] = *reinterpret_cast<float2*>(&cc);
        }
    }
    float s = 0; for (int j = 0; j < 16; ++j) s += c[j].x + c[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class K> float run(K k, float* out, int blocks, int threads, int iters) {
    k<<<blocks, threads>>>(out, iters); cudaDeviceSynchronize();
    cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
    cudaEventRecord(s); k<<<blocks, threads>>>(out, iters); cudaEventRecord(e); cudaEventSynchronize(e);
    float ms; cudaEventElapsedTime(&ms, s, e); return ms;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    const int iters = 20000, blocks = 148 * 4, threads = 256;
    float ms = run(mma_tf32, out, blocks, threads, iters);
    double fl = 2.0 * 16 * 8 * 8 * 8.0 * iters * blocks * (threads / 32);
    printf("mma.sync m16n8k8 tf32 : %.2f ms  %.1f TFLOP/s\n", ms, fl / ms / 1e9);
    ms = run(mma_bf16, out, blocks, threads, iters);
    fl = 2.0 * 16 * 8 * 16 * 8.0 * iters * blocks * (threads / 32);
    printf("mma.sync m16n8k16 bf16: %.2f ms  %.1f TFLOP/s\n", ms, fl / ms / 1e9);
    ms = run(ffma1, out, blocks, threads, iters);
    fl = 2.0 * 32 * (double)iters * blocks * threads;
    printf("FFMA                  : %.2f ms  %.1f TFLOP/s\n", ms, fl / ms / 1e9);
    ms = run(ffma2, out, blocks, threads, iters);
    printf("FFMA2 (fma.rn.f32x2)  : %.2f ms  %.1f TFLOP/s\n", ms, fl / ms / 1e9);
    cudaError_t err = cudaGetLastError(); printf("status: %s\n", cudaGetErrorString(err));
    return 0;
}
