// 3x3 convolution, TF 'SAME' padding, + bias [+ leaky] [+ residual] on the CUDA cores (exact fp32).
//
// Replaces tf.layers.Conv2D(filters,(3,3),strides,'same',dilation_rate) + tf.nn.leaky_relu
// (+ the residual adds) of the reference: modules.py:62-67 (pyramid), 266-277 (estimator),
// 306-326 (context).  This is the any-shape path: the first pyramid conv (Cin=3, K=27), the
// two-channel flow heads (N=2), the stride-2 pyramid convs, use_dc=True channel counts, and the
// numerical yard-stick for the tcgen05 implicit-GEMM path in conv_tc.cu.
//
//  * conv3x3_tiled_kernel: implicit GEMM, 128 pixels x 64 output channels per CTA, K = 9 taps x
//    8-channel slices staged in shared memory, 8x4 register tile per thread.
//  * conv3x3_thin_kernel<COUT>: one thread per output pixel for COUT in {2,16} (memory-bound
//    layers), weights broadcast from shared memory.
#include "common.cuh"

namespace pwc {

struct ConvParams {
    const float* x; const float* w; const float* bias; const float* res; float* y;
    int x_cs, res_cs, y_cs;
    int B, H, W, Cin, Cout, OH, OW;
    int stride, dil, pad_t, pad_l;
    float alpha;
    int vec_x;   // x loads may use float4 (16B-aligned base, x_cs % 4 == 0)
};

constexpr int CT_BM = 128, CT_BN = 64, CT_BK = 8, CT_THREADS = 256;
constexpr int CT_TW = 16, CT_TH = 8;   // pixel tile 8 rows x 16 cols

__global__ void __launch_bounds__(CT_THREADS) conv3x3_tiled_kernel(const ConvParams p) {
    __shared__ __align__(16) float As[CT_BK][CT_BM + 4];
    __shared__ __align__(16) float Bs[CT_BK][CT_BN];

    const int tid = threadIdx.x;
    const int tiles_x = (p.OW + CT_TW - 1) / CT_TW;
    const int tile_y = blockIdx.x / tiles_x, tile_x = blockIdx.x - tile_y * tiles_x;
    const int b = blockIdx.z;
    const int n0 = blockIdx.y * CT_BN;
    const float* xb = p.x + (size_t)b * p.H * p.W * p.x_cs;

    // A-load role: pixel m = tid & 127, channel half kg = tid >> 7 (4 channels)
    const int lm = tid & (CT_BM - 1), lkg = tid >> 7;
    const int l_oy = tile_y * CT_TH + (lm >> 4), l_ox = tile_x * CT_TW + (lm & 15);
    // B-load role: 8 x 64 floats = 512 -> 2 per thread
    // compute role: thread -> 8 pixels (tm*8..) x 4 couts (tn*4..)
    const int tn = tid & 15, tm = tid >> 4;

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int tap = 0; tap < 9; ++tap) {
        const int ky = tap / 3, kx = tap - ky * 3;
        const int iy = l_oy * p.stride - p.pad_t + ky * p.dil;
        const int ix = l_ox * p.stride - p.pad_l + kx * p.dil;
        const bool inb = (l_oy < p.OH) && (l_ox < p.OW) && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
        const float* xp = xb + ((size_t)iy * p.W + ix) * p.x_cs;
        const float* wt = p.w + (size_t)tap * p.Cin * p.Cout;
        for (int c0 = 0; c0 < p.Cin; c0 += CT_BK) {
            // ---- stage A (transposed to [k][m]) and B
            {
                const int c = c0 + 4 * lkg;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (inb) {
                    if (p.vec_x && c + 3 < p.Cin) v = ldg4(xp + c);
                    else {
                        if (c + 0 < p.Cin) v.x = __ldg(xp + c + 0);
                        if (c + 1 < p.Cin) v.y = __ldg(xp + c + 1);
                        if (c + 2 < p.Cin) v.z = __ldg(xp + c + 2);
                        if (c + 3 < p.Cin) v.w = __ldg(xp + c + 3);
                    }
                }
                As[4 * lkg + 0][lm] = v.x; As[4 * lkg + 1][lm] = v.y;
                As[4 * lkg + 2][lm] = v.z; As[4 * lkg + 3][lm] = v.w;
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int e = tid + r * CT_THREADS;
                    const int k = e >> 6, n = e & 63;
                    float wv = 0.f;
                    if (c0 + k < p.Cin && n0 + n < p.Cout) wv = __ldg(wt + (size_t)(c0 + k) * p.Cout + n0 + n);
                    Bs[k][n] = wv;
                }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < CT_BK; ++k) {
                const float4 a0 = *reinterpret_cast<const float4*>(&As[k][tm * 8]);
                const float4 a1 = *reinterpret_cast<const float4*>(&As[k][tm * 8 + 4]);
                const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tn * 4]);
                const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
    // ---- epilogue: bias, leaky, residual
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = tm * 8 + i;
        const int oy = tile_y * CT_TH + (m >> 4), ox = tile_x * CT_TW + (m & 15);
        if (oy >= p.OH || ox >= p.OW) continue;
        const size_t pix = ((size_t)b * p.OH + oy) * p.OW + ox;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tn * 4 + j;
            if (n >= p.Cout) continue;
            float v = acc[i][j] + __ldg(p.bias + n);
            v = leaky(v, p.alpha);
            if (p.res) v += __ldg(p.res + pix * p.res_cs + n);
            p.y[pix * p.y_cs + n] = v;
        }
    }
}

constexpr int THIN_MAX_W = 8192;   // floats of weights staged in shared memory

template <int COUT>
__global__ void __launch_bounds__(128) conv3x3_thin_kernel(const ConvParams p) {
    __shared__ __align__(16) float ws[THIN_MAX_W];
    const int nw = 9 * p.Cin * COUT;
    for (int e = threadIdx.x; e < nw; e += blockDim.x) ws[e] = __ldg(p.w + e);
    __syncthreads();
    const size_t total = (size_t)p.B * p.OH * p.OW;
    for (size_t pix = blockIdx.x * (size_t)blockDim.x + threadIdx.x; pix < total; pix += (size_t)gridDim.x * blockDim.x) {
        const int ox = pix % p.OW; const size_t row = pix / p.OW;
        const int oy = row % p.OH; const int b = row / p.OH;
        const float* xb = p.x + (size_t)b * p.H * p.W * p.x_cs;
        float acc[COUT];
#pragma unroll
        for (int j = 0; j < COUT; ++j) acc[j] = 0.f;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
            const int ky = tap / 3, kx = tap - ky * 3;
            const int iy = oy * p.stride - p.pad_t + ky * p.dil;
            const int ix = ox * p.stride - p.pad_l + kx * p.dil;
            if (iy < 0 || iy >= p.H || ix < 0 || ix >= p.W) continue;
            const float* xp = xb + ((size_t)iy * p.W + ix) * p.x_cs;
            const float* wt = ws + tap * p.Cin * COUT;
            int c = 0;
            if (p.vec_x) {
                for (; c + 3 < p.Cin; c += 4) {
                    const float4 v = ldg4(xp + c);
                    const float xv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q)
#pragma unroll
                        for (int j = 0; j < COUT; ++j) acc[j] = fmaf(xv[q], wt[(c + q) * COUT + j], acc[j]);
                }
            }
            for (; c < p.Cin; ++c) {
                const float v = __ldg(xp + c);
#pragma unroll
                for (int j = 0; j < COUT; ++j) acc[j] = fmaf(v, wt[c * COUT + j], acc[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < COUT; ++j) {
            float v = acc[j] + __ldg(p.bias + j);
            v = leaky(v, p.alpha);
            if (p.res) v += __ldg(p.res + pix * p.res_cs + j);
            p.y[pix * p.y_cs + j] = v;
        }
    }
}

// TF SAME padding (before) for one dim: out = ceil(in/s); total = max((out-1)s + 2d + 1 - in, 0).
static inline void same_pad(int in, int stride, int dil, int* out, int* before) {
    *out = (in + stride - 1) / stride;
    int total = (*out - 1) * stride + 2 * dil + 1 - in;
    if (total < 0) total = 0;
    *before = total / 2;
}

}  // namespace pwc

extern "C" int pwc_conv3x3_fwd(const float* x, int x_cs, const float* w_hwio, const float* bias,
                               const float* residual, int res_cs, float* y, int y_cs,
                               int B, int H, int W, int Cin, int Cout, int stride, int dilation,
                               float alpha, void* stream) {
    using namespace pwc;
    PWC_REQUIRE(x && w_hwio && bias && y, PWC_E_BADARG, "conv3x3: null pointer");
    PWC_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, PWC_E_BADARG, "conv3x3: bad dims");
    PWC_REQUIRE(stride >= 1 && dilation >= 1, PWC_E_BADARG, "conv3x3: stride/dilation must be >= 1");
    PWC_REQUIRE(x_cs >= Cin && y_cs >= Cout, PWC_E_BADARG, "conv3x3: channel stride smaller than channel count");
    PWC_REQUIRE(B <= 65535, PWC_E_BADARG, "conv3x3: batch > 65535");
    ConvParams p{};
    p.x = x; p.w = w_hwio; p.bias = bias; p.res = residual; p.y = y;
    p.x_cs = x_cs; p.res_cs = res_cs; p.y_cs = y_cs;
    p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
    p.stride = stride; p.dil = dilation; p.alpha = alpha;
    same_pad(H, stride, dilation, &p.OH, &p.pad_t);
    same_pad(W, stride, dilation, &p.OW, &p.pad_l);
    p.vec_x = aligned16(x) && (x_cs % 4 == 0);
    cudaStream_t st = (cudaStream_t)stream;
    const bool thin = (Cout == 2 || Cout == 16) && (9 * Cin * Cout <= THIN_MAX_W) && Cin <= 64;
    if (thin) {
        const size_t total = (size_t)B * p.OH * p.OW;
        const int blocks = (int)((total + 127) / 128 < (size_t)148 * 16 ? (total + 127) / 128 : (size_t)148 * 16);
        if (Cout == 2) conv3x3_thin_kernel<2><<<blocks, 128, 0, st>>>(p);
        else conv3x3_thin_kernel<16><<<blocks, 128, 0, st>>>(p);
        PWC_CHECK_LAUNCH("conv3x3_thin_kernel");
    } else {
        const int tiles = ((p.OW + CT_TW - 1) / CT_TW) * ((p.OH + CT_TH - 1) / CT_TH);
        dim3 grid(tiles, (Cout + CT_BN - 1) / CT_BN, B);
        conv3x3_tiled_kernel<<<grid, CT_THREADS, 0, st>>>(p);
        PWC_CHECK_LAUNCH("conv3x3_tiled_kernel");
    }
    return 0;
}
