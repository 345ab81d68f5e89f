// First pyramid convolution: 3 -> 16 channels, 3x3, stride 2, TF 'SAME' padding, + bias + leaky (modules.py:62-63 with
// l = 0: tf.layers.Conv2D(16, (3,3), (2,2), 'same') + tf.nn.leaky_relu(0.1)) -- on the CUDA cores, HBM-bound.
//
// K = 27 is too small for the tensor cores; the layer is 432 FMA per output pixel against 12 input bytes (uint8) and 64
// output bytes: the roofline is the 22 MB (uint8 images) or 88 MB (fp32) read + the 117 MB write at B = 8 x 448 x 1024.
// The round-1 kernel (one thread per output pixel, taps gathered from global memory, weights in shared memory, 64-byte
// strided stores) ran at 185 us.  Here:
//   * CTA = 8 x 64 output pixels; the 17 x 129 x 3 input patch is staged in shared memory once (coalesced 4-byte loads;
//     uint8 inputs go through the 256-entry table float32(float64(v) / 255.0), i.e. the reference's `images / 255.0`
//     feed, test.py:31-33, train.py:122 -- the bytes themselves cross PCIe and HBM, nothing is expanded in memory);
//   * a thread owns four output pixels x 16 channels (64 accumulators); the 27 x 16 weights sit in __constant__ memory:
//     one constant load per weight feeds four FFMAs, one LDS.32 per input value feeds 16;
//   * the 16 results of a pixel go through a shared-memory transpose so that a warp's float4 stores cover 512 contiguous
//     bytes (16-byte-per-lane stores at a 64-byte stride cost 2.2x, profiles/r02_tmem_ld_bench.log).
// Exact fp32 (same operation order per output: taps row-major, channels inner), the yard-stick class of conv_direct.cu.
// The one stateful entry point of the library: the layer's 448 weights are uploaded into a module-global __constant__
// buffer in stream order before every launch; concurrent calls on different streams with different weights must be
// serialised by the caller (one model per process and GPU, as everywhere in this package).
#include "common.cuh"

namespace pwc {

constexpr int F_TH = 8, F_TW = 64, F_CO = 16;
constexpr int F_PH = 2 * F_TH + 1, F_PW = 2 * F_TW + 1;       // 17 x 129 input pixels
constexpr int F_THREADS = 128;                                // 4 output pixels x 16 channels per thread
constexpr int F_PPT = 4;                                      // (ncu, round 2: with 2 pixels per thread the kernel was instruction-
                                                              // bound -- one LDC per weight per 2 FFMAs, issue slots 76 % busy, FMA pipe 39 %)
constexpr int F_ROW = (F_PW * 3 + 3) / 4 * 4;                 // patch row pitch in floats: 388 (97 whole 4-byte-pixel words)
constexpr int F_PATCH = F_PH * F_ROW;                         // floats
constexpr int F_OPITCH = F_CO + 4;                            // staging row pitch in floats (conflict-free float4 stores)

__constant__ float c_first_w[27 * F_CO + F_CO];               // HWIO kernel (ky, kx, ci, co) then the bias

struct FirstParams {
    const void* x; float* y; const float* lut;
    int y_cs, B, H, W, OH, OW;
    float alpha;
};

template <bool U8>
__global__ void __launch_bounds__(F_THREADS) conv_first_kernel(const FirstParams p) {
    // the output staging area re-uses the patch's shared memory (40 KB; the patch needs 26 KB)
    __shared__ __align__(16) float stage[F_TH * F_TW * F_OPITCH];
    __shared__ float lut[256];
    float* patch = stage;
    static_assert(F_PATCH <= F_TH * F_TW * F_OPITCH, "patch must fit in the staging area");
    const int tid = threadIdx.x;
    const int tiles_x = (p.OW + F_TW - 1) / F_TW, tiles_y = (p.OH + F_TH - 1) / F_TH;
    const int tile = blockIdx.x;
    const int b = tile / (tiles_x * tiles_y), r = tile - b * tiles_x * tiles_y;
    const int oy0 = (r / tiles_x) * F_TH, ox0 = (r % tiles_x) * F_TW;
    const int iy0 = 2 * oy0, ix0 = 2 * ox0;                    // SAME padding of an even size with stride 2: 0 before, 1 after
    if (U8) {
        lut[tid] = __ldg(p.lut + tid);
        lut[tid + F_THREADS] = __ldg(p.lut + tid + F_THREADS);
        __syncthreads();
    }
    // ---- stage the input patch: rows iy0 .. iy0+16, columns ix0 .. ix0+128, 3 channels; zero outside the image
    const int row_elems = F_PW * 3;                            // 387 values per patch row, stored at a pitch of F_ROW
    if (U8) {
        const uint8_t* xb = static_cast<const uint8_t*>(p.x) + (size_t)b * p.H * p.W * 3;
        constexpr int words = F_ROW / 4;                       // 97 4-byte words per row (row starts are 4-byte aligned: W % 4 == 0)
        // all of a thread's 4-byte loads are issued before the first table lookup (13 independent loads in flight: as a
        // plain loop the staging phase was 13 serialised global-memory round trips per CTA)
        constexpr int NIT = (F_PH * words + F_THREADS - 1) / F_THREADS;
        uint32_t vv[NIT];
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
            const int e = tid + i * F_THREADS;
            const int ry = e / words, w4 = e - ry * words;
            const int iy = iy0 + ry;
            const long long off = ((long long)iy * p.W + ix0) * 3 + 4 * w4;       // byte offset inside the image
            const long long left = ((long long)iy + 1) * p.W * 3 - off;           // bytes up to the end of the image row
            uint32_t v = 0;
            if (e < F_PH * words && iy < p.H && left > 0) {
                if (left >= 4) v = __ldg(reinterpret_cast<const uint32_t*>(xb + off));
                else {
                    for (int k = 0; k < (int)left; ++k) v |= (uint32_t)xb[off + k] << (8 * k);
                }
            }
            vv[i] = v;
        }
        // bytes beyond the image are 0 and the table maps 0 to 0.0f: the zero padding needs no branches
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
            const int e = tid + i * F_THREADS;
            const int ry = e / words, w4 = e - ry * words;
            const uint32_t v = vv[i];
            if (e < F_PH * words)
                *reinterpret_cast<float4*>(patch + ry * F_ROW + 4 * w4) =
                    make_float4(lut[v & 0xFF], lut[(v >> 8) & 0xFF], lut[(v >> 16) & 0xFF], lut[v >> 24]);
        }
    } else {
        const float* xb = static_cast<const float*>(p.x) + (size_t)b * p.H * p.W * 3;
        for (int e = tid; e < F_PH * row_elems; e += F_THREADS) {
            const int ry = e / row_elems, col = e - ry * row_elems;
            const int iy = iy0 + ry;
            const long long off = ((long long)iy * p.W + ix0) * 3 + col;
            patch[ry * F_ROW + col] = (iy < p.H && off < ((long long)iy + 1) * p.W * 3) ? __ldg(xb + off) : 0.f;
        }
    }
    __syncthreads();
    // ---- four output pixels per thread: (oy + 2 h, ox), h = 0..3; lanes = consecutive ox
    const int lx = tid & (F_TW - 1), ly = tid >> 6;            // 64 x 2
    float acc[F_PPT][F_CO];
#pragma unroll
    for (int h = 0; h < F_PPT; ++h)
#pragma unroll
        for (int j = 0; j < F_CO; ++j) acc[h][j] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
                const int widx = ((ky * 3 + kx) * 3 + ci) * F_CO;
                float xv[F_PPT];
#pragma unroll
                for (int h = 0; h < F_PPT; ++h) xv[h] = patch[(2 * (ly + 2 * h) + ky) * F_ROW + (2 * lx + kx) * 3 + ci];
#pragma unroll
                for (int j = 0; j < F_CO; ++j) {
                    const float wv = c_first_w[widx + j];
#pragma unroll
                    for (int h = 0; h < F_PPT; ++h) acc[h][j] = fmaf(xv[h], wv, acc[h][j]);
                }
            }
    // ---- bias, leaky, transpose through shared memory, coalesced float4 stores
    __syncthreads();                                           // every thread is done with the patch
#pragma unroll
    for (int h = 0; h < F_PPT; ++h) {
        float* s = stage + ((ly + 2 * h) * F_TW + lx) * F_OPITCH;
#pragma unroll
        for (int j = 0; j < F_CO; j += 4) {
            float4 v;
            v.x = leaky(acc[h][j] + c_first_w[27 * F_CO + j], p.alpha);
            v.y = leaky(acc[h][j + 1] + c_first_w[27 * F_CO + j + 1], p.alpha);
            v.z = leaky(acc[h][j + 2] + c_first_w[27 * F_CO + j + 2], p.alpha);
            v.w = leaky(acc[h][j + 3] + c_first_w[27 * F_CO + j + 3], p.alpha);
            *reinterpret_cast<float4*>(s + j) = v;
        }
    }
    __syncthreads();
    const bool vec = (p.y_cs & 3) == 0 && aligned16(p.y);
    // unit = (pixel, 4-channel chunk): 512 pixels x 4 chunks; consecutive threads -> consecutive chunks -> contiguous bytes
    for (int u = tid; u < F_TH * F_TW * 4; u += F_THREADS) {
        const int pix = u >> 2, ch = (u & 3) * 4;
        const int oy = oy0 + pix / F_TW, ox = ox0 + (pix & (F_TW - 1));
        if (oy >= p.OH || ox >= p.OW) continue;
        const float4 v = *reinterpret_cast<const float4*>(stage + pix * F_OPITCH + ch);
        float* dst = p.y + (((size_t)b * p.OH + oy) * p.OW + ox) * p.y_cs + ch;
        if (vec) *reinterpret_cast<float4*>(dst) = v;
        else { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w; }
    }
}

}  // namespace pwc

extern "C" int pwc_conv_first_fwd(const void* x, int x_is_u8, const float* lut256, const float* w_hwio, const float* bias,
                                  float* y, int y_cs, int B, int H, int W, float alpha, void* stream) {
    using namespace pwc;
    PWC_REQUIRE(x && w_hwio && bias && y, PWC_E_BADARG, "conv_first: null pointer");
    PWC_REQUIRE(B > 0 && H > 0 && W > 0 && y_cs >= F_CO, PWC_E_BADARG, "conv_first: bad dims");
    PWC_REQUIRE((H & 1) == 0 && (W & 3) == 0, PWC_E_BADARG, "conv_first: H must be even and W a multiple of 4 (the network needs /64)");
    PWC_REQUIRE(!x_is_u8 || lut256, PWC_E_BADARG, "conv_first: uint8 input needs the 256-entry table");
    PWC_REQUIRE((reinterpret_cast<uintptr_t>(x) & 3u) == 0, PWC_E_ALIGN, "conv_first: x must be 4-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    // weights + bias -> constant memory (device-to-device, stream-ordered: captured as a memcpy node in CUDA graphs, so a
    // replay picks up the current weights)
    cudaError_t e = cudaMemcpyToSymbolAsync(c_first_w, w_hwio, 27 * F_CO * sizeof(float), 0, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess)
        e = cudaMemcpyToSymbolAsync(c_first_w, bias, F_CO * sizeof(float), 27 * F_CO * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) { set_error("conv_first: constant upload: %s", cudaGetErrorString(e)); return (int)e; }
    FirstParams p{};
    p.x = x; p.y = y; p.lut = lut256; p.y_cs = y_cs; p.B = B; p.H = H; p.W = W; p.OH = H / 2; p.OW = W / 2; p.alpha = alpha;
    const long long tiles = (long long)B * ((p.OW + F_TW - 1) / F_TW) * ((p.OH + F_TH - 1) / F_TH);
    PWC_REQUIRE(tiles < (1ll << 31), PWC_E_BADARG, "conv_first: too many tiles");
    if (x_is_u8) conv_first_kernel<true><<<(int)tiles, F_THREADS, 0, st>>>(p);
    else conv_first_kernel<false><<<(int)tiles, F_THREADS, 0, st>>>(p);
    PWC_CHECK_LAUNCH("conv_first_kernel");
    return 0;
}
