// Backward kernels of the PWC-Net hot path (training step, reference train.py:65-92) -- exact fp32, sm_100a.
//
// TensorFlow 1.8 autodiff built these node groups for the reference; each kernel replaces one group:
//   pwc_conv3x3_dgrad          <- Conv2DBackpropInput (+ LeakyReluGrad of the producing layer in the epilogue)
//   pwc_conv3x3_wgrad          <- Conv2DBackpropFilter + BiasAddGrad (+ the Cin un-permutation of concat layers)
//   pwc_leaky_bwd              <- LeakyReluGrad for activations with several consumers
//   pwc_add_strided            <- AddN of gradient contributions (residual adds modules.py:275-277,326, concat slots)
//   pwc_cost_volume_bwd        <- gradient of the 81 x (pad, pad, mul, crop, mean) + stack + leaky (modules.py:164-204)
//   pwc_warp_bwd               <- ScatterNd x 4 + the bilinear-weight gradient w.r.t. the flow (modules.py:99-137)
//   pwc_resize_bilinear_bwd    <- ResizeBilinearGrad (modules.py:283-284)
//   pwc_lploss_level_bwd       <- gradient of L1loss/L2loss(resize_nearest(gt/20), fs) (losses.py:4-8,20-29)
//   pwc_adam_step              <- l2_loss regulariser gradient (train.py:74) + tf.train.AdamOptimizer (train.py:89)
//   pwc_sumsq                  <- tf.nn.l2_loss (train.py:74) for the reported total loss
//   pwc_permute_cin            <- rebuilds the internal-channel-order kernels of the concat layers after an update
//
// Gradient buffers mirror the activation buffers (same NHWC shapes and channel strides).  A buffer holds the
// gradient w.r.t. the PRE-activation output of its layer by the time that layer's dgrad/wgrad run.
#include "common.cuh"
#include <cstdlib>
#include <cstring>

namespace pwc {

static inline void same_pad_b(int in, int stride, int dil, int* out, int* before) {
    *out = (in + stride - 1) / stride;
    int total = (*out - 1) * stride + 2 * dil + 1 - in;
    if (total < 0) total = 0;
    *before = total / 2;
}

static inline int grid_for(size_t total, int threads, int max_blocks = 148 * 16) {
    size_t b = (total + threads - 1) / threads;
    if (b < 1) b = 1;
    return (int)(b < (size_t)max_blocks ? b : (size_t)max_blocks);
}

// ------------------------------------------------------------------------------------------------ dgrad
struct DgradParams {
    const float* dy; const float* w; const float* mask; float* dx;
    int dy_cs, mask_cs, dx_cs;
    int B, H, W, Cin, Cout, OH, OW;
    int stride, dil, pad_t, pad_l;
    float mask_alpha;
    int accumulate, vec_dy;
};

constexpr int DG_BM = 128, DG_BN = 64, DG_BK = 8, DG_THREADS = 256, DG_TW = 16, DG_TH = 8;

// Implicit GEMM over INPUT pixels: dx[b,iy,ix,ci] = sum_{tap,co} dy[b,oy,ox,co] * w[tap,ci,co] with
// oy*stride - pad_t + ky*dil == iy (same for x).  128 input pixels x 64 input channels per CTA, K = 9 taps x
// 8-channel slices of Cout, 8x4 register tile per thread.
__global__ void __launch_bounds__(DG_THREADS) conv3x3_dgrad_kernel(const DgradParams p) {
    __shared__ __align__(16) float As[DG_BK][DG_BM + 4];
    __shared__ __align__(16) float Bs[DG_BK][DG_BN + 4];

    // For stride s the input pixels are tiled per parity class (iy % s, ix % s): every pixel of a tile then sees the
    // same subset of taps (those with (iy + pad_t - ky*dil) % s == 0), the others are skipped instead of multiplied by 0.
    const int tid = threadIdx.x;
    const int sub = p.stride;
    const int n_tiles_n = (p.Cin + DG_BN - 1) / DG_BN;
    const int par = blockIdx.y / n_tiles_n, ry = par / sub, rx = par - ry * sub;
    const int Hs = (p.H - ry + sub - 1) / sub, Ws = (p.W - rx + sub - 1) / sub;   // pixels of this parity class
    const int tiles_x = ((p.W + sub - 1) / sub + DG_TW - 1) / DG_TW;
    const int tile_y = blockIdx.x / tiles_x, tile_x = blockIdx.x - tile_y * tiles_x;
    const int b = blockIdx.z;
    const int n0 = (blockIdx.y - par * n_tiles_n) * DG_BN;
    const float* dyb = p.dy + (size_t)b * p.OH * p.OW * p.dy_cs;

    const int lm = tid & (DG_BM - 1), lkg = tid >> 7;
    const int l_sy = tile_y * DG_TH + (lm >> 4), l_sx = tile_x * DG_TW + (lm & 15);
    const int l_iy = l_sy * sub + ry, l_ix = l_sx * sub + rx;
    const bool l_ok = l_sy < Hs && l_sx < Ws;
    const int tn = tid & 15, tm = tid >> 4;
    // weight-load role (threads 0..127): input channel n = tid >> 1, four output channels 4*(tid & 1)..
    const int wn = tid >> 1, wh = tid & 1;

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int tap = 0; tap < 9; ++tap) {
        const int ky = tap / 3, kx = tap - ky * 3;
        {   // tap parity is uniform over the CTA: skip taps that never hit an output pixel of this class
            const int py2 = ((ry + p.pad_t - ky * p.dil) % sub + sub) % sub, px2 = ((rx + p.pad_l - kx * p.dil) % sub + sub) % sub;
            if (py2 != 0 || px2 != 0) continue;
        }
        const int ty = l_iy + p.pad_t - ky * p.dil, tx = l_ix + p.pad_l - kx * p.dil;
        bool inb = l_ok && ty >= 0 && tx >= 0;
        const int oy = ty / p.stride, ox = tx / p.stride;
        inb = inb && oy < p.OH && ox < p.OW;
        const float* dyp = dyb + ((size_t)oy * p.OW + ox) * p.dy_cs;
        const float* wt = p.w + (size_t)tap * p.Cin * p.Cout;
        for (int c0 = 0; c0 < p.Cout; c0 += DG_BK) {
            {
                const int c = c0 + 4 * lkg;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (inb) {
                    if (p.vec_dy && c + 3 < p.Cout) v = ldg4(dyp + c);
                    else {
                        if (c + 0 < p.Cout) v.x = __ldg(dyp + c + 0);
                        if (c + 1 < p.Cout) v.y = __ldg(dyp + c + 1);
                        if (c + 2 < p.Cout) v.z = __ldg(dyp + c + 2);
                        if (c + 3 < p.Cout) v.w = __ldg(dyp + c + 3);
                    }
                }
                As[4 * lkg + 0][lm] = v.x; As[4 * lkg + 1][lm] = v.y;
                As[4 * lkg + 2][lm] = v.z; As[4 * lkg + 3][lm] = v.w;
                if (tid < 128) {
                    const int ci = n0 + wn;
                    const float* wp = wt + (size_t)ci * p.Cout + c0 + 4 * wh;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float wv = 0.f;
                        if (ci < p.Cin && c0 + 4 * wh + q < p.Cout) wv = __ldg(wp + q);
                        Bs[4 * wh + q][wn] = wv;
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < DG_BK; ++k) {
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
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = tm * 8 + i;
        const int sy = tile_y * DG_TH + (m >> 4), sx = tile_x * DG_TW + (m & 15);
        if (sy >= Hs || sx >= Ws) continue;
        const int iy = sy * sub + ry, ix = sx * sub + rx;
        const size_t pix = ((size_t)b * p.H + iy) * p.W + ix;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tn * 4 + j;
            if (n >= p.Cin) continue;
            float v = acc[i][j];
            if (p.mask) v *= (__ldg(p.mask + pix * p.mask_cs + n) > 0.f) ? 1.f : p.mask_alpha;
            float* d = p.dx + pix * p.dx_cs + n;
            if (p.accumulate) v += *d;
            *d = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------ wgrad
struct WgradParams {
    const float* x; const float* dy; float* dw; float* db; const int* cin_map;
    int x_cs, dy_cs;
    int B, H, W, Cin, Cout, OH, OW;
    int stride, dil, pad_t, pad_l;
    int dw_cin;          // number of input channels of the destination kernel (reference order)
    int ci_tiles;
    long long total;     // output pixels B*OH*OW
    long long chunk;     // output pixels per CTA (multiple of WG_BK)
    int vec_x, vec_dy;
};

constexpr int WG_BK = 8, WG_THREADS = 256;

// dw[tap,ci,co] += sum_{output pixels} x[pixel shifted by tap, ci] * dy[pixel, co]:  per tap a GEMM with
// M = Cin, N = Cout, K = pixels.  CTA = (pixel chunk, TM x TN tile of (ci,co), tap); both operands are staged in
// their natural channel-contiguous layout (8 pixels per stage, the next stage is fetched into registers while the
// current one is multiplied), (TM/16) x (TN/16) register tile per thread (8x8 for the 128-channel layers: 64 FMAs per
// four LDS.128), atomic accumulation into dw.  CTAs of tap 0 / ci-tile 0 also reduce db[co] += sum dy[pixel, co].
template <int TM, int TN>
__global__ void __launch_bounds__(WG_THREADS, 2) conv3x3_wgrad_kernel(const WgradParams p) {
    constexpr int BK = (TM == 64 && TN == 64) ? 16 : WG_BK;   // pixels per stage: every thread loads one float4 per operand
    constexpr int RM = TM / 16, RN = TN / 16;         // register tile (4 or 8 per dimension), split in halves of 4
    __shared__ __align__(16) float As[BK][TM];
    __shared__ __align__(16) float Bs[BK][TN];

    const int tid = threadIdx.x;
    const int tap = blockIdx.z, ky = tap / 3, kx = tap - ky * 3;
    const int ci0 = (blockIdx.y % p.ci_tiles) * TM, co0 = (blockIdx.y / p.ci_tiles) * TN;
    const long long p_begin = (long long)blockIdx.x * p.chunk;
    const long long p_end = p_begin + p.chunk < p.total ? p_begin + p.chunk : p.total;
    // loader roles: pixel lk of the stage, four channels starting at la (x) / lb (dy)
    const int lk = tid / (TM / 4) < BK ? tid / (TM / 4) : 0, la = (tid % (TM / 4)) * 4;
    const bool a_role = tid < BK * (TM / 4);
    const int lkb = tid / (TN / 4) < BK ? tid / (TN / 4) : 0, lb = (tid % (TN / 4)) * 4;
    const bool b_role = tid < BK * (TN / 4);
    const int tm = tid >> 4, tn = tid & 15;
    const bool do_bias = p.db != nullptr && tap == 0 && ci0 == 0;

    float acc[RM][RN];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) acc[i][j] = 0.f;
    float bsum = 0.f;

    auto fetch = [&](long long p0, float4& av, float4& bv) {
        av = make_float4(0.f, 0.f, 0.f, 0.f); bv = av;
        if (b_role) {
            const long long pix = p0 + lkb;
            if (pix < p_end) {
                const float* dyp = p.dy + (size_t)pix * p.dy_cs;
                const int co = co0 + lb;
                if (p.vec_dy && co + 3 < p.Cout) bv = ldg4(dyp + co);
                else {
                    if (co + 0 < p.Cout) bv.x = __ldg(dyp + co + 0);
                    if (co + 1 < p.Cout) bv.y = __ldg(dyp + co + 1);
                    if (co + 2 < p.Cout) bv.z = __ldg(dyp + co + 2);
                    if (co + 3 < p.Cout) bv.w = __ldg(dyp + co + 3);
                }
            }
        }
        if (a_role) {
            const long long pix = p0 + lk;
            if (pix < p_end) {
                const int ox = (int)(pix % p.OW); const long long row = pix / p.OW;
                const int oy = (int)(row % p.OH); const int b = (int)(row / p.OH);
                const int iy = oy * p.stride - p.pad_t + ky * p.dil, ix = ox * p.stride - p.pad_l + kx * p.dil;
                if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
                    const float* xp = p.x + (((size_t)b * p.H + iy) * p.W + ix) * p.x_cs;
                    const int ci = ci0 + la;
                    if (p.vec_x && ci + 3 < p.Cin) av = ldg4(xp + ci);
                    else {
                        if (ci + 0 < p.Cin) av.x = __ldg(xp + ci + 0);
                        if (ci + 1 < p.Cin) av.y = __ldg(xp + ci + 1);
                        if (ci + 2 < p.Cin) av.z = __ldg(xp + ci + 2);
                        if (ci + 3 < p.Cin) av.w = __ldg(xp + ci + 3);
                    }
                }
            }
        }
    };

    float4 av, bv;
    fetch(p_begin, av, bv);
    for (long long p0 = p_begin; p0 < p_end; p0 += BK) {
        if (a_role) *reinterpret_cast<float4*>(&As[lk][la]) = av;
        if (b_role) *reinterpret_cast<float4*>(&Bs[lkb][lb]) = bv;
        __syncthreads();
        if (p0 + BK < p_end) fetch(p0 + BK, av, bv);     // in flight during the multiply
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[RM], bb[RN];
#pragma unroll
            for (int h = 0; h < RM / 4; ++h) {
                const float4 t4 = *reinterpret_cast<const float4*>(&As[k][h * (TM / 2) + tm * 4]);
                a[4 * h] = t4.x; a[4 * h + 1] = t4.y; a[4 * h + 2] = t4.z; a[4 * h + 3] = t4.w;
            }
#pragma unroll
            for (int h = 0; h < RN / 4; ++h) {
                const float4 t4 = *reinterpret_cast<const float4*>(&Bs[k][h * (TN / 2) + tn * 4]);
                bb[4 * h] = t4.x; bb[4 * h + 1] = t4.y; bb[4 * h + 2] = t4.z; bb[4 * h + 3] = t4.w;
            }
#pragma unroll
            for (int i = 0; i < RM; ++i)
#pragma unroll
                for (int j = 0; j < RN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        if (do_bias && tid < TN) {
#pragma unroll
            for (int k = 0; k < BK; ++k) bsum += Bs[k][tid];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < RM; ++i) {
        const int ci = ci0 + (i >> 2) * (TM / 2) + tm * 4 + (i & 3);
        if (ci >= p.Cin) continue;
        const int dst = p.cin_map ? __ldg(p.cin_map + ci) : ci;
        if (dst < 0) continue;
#pragma unroll
        for (int j = 0; j < RN; ++j) {
            const int co = co0 + (j >> 2) * (TN / 2) + tn * 4 + (j & 3);
            if (co < p.Cout) atomicAdd(p.dw + ((size_t)tap * p.dw_cin + dst) * p.Cout + co, acc[i][j]);
        }
    }
    if (do_bias && tid < TN && co0 + tid < p.Cout) atomicAdd(p.db + co0 + tid, bsum);
}

// Small-channel wgrad (Cin <= 32 and Cout <= 32, dilation 1: the first two pyramid levels and the 2-channel flow
// heads).  These layers are memory/latency-bound, so one persistent CTA handles ALL nine taps of a 4 x 32 output tile
// from one shared-memory patch of x and one tile of dy (each input byte is read once instead of nine times), keeps
// its (tap, 4 ci, 4 co) register tiles across all of its tiles and issues its atomics once at the end.
__device__ __forceinline__ void cp_async16(float* dst, const float* src, bool ok) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const int sz = ok ? 16 : 0;                          // src-size 0: nothing is read, the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async4(float* dst, const float* src, bool ok) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const int sz = ok ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int WS_TW = 32, WS_TH = 4, WS_PIX = WS_TW * WS_TH, WS_MAXT = 288;
// Work items = (tap, TI input channels, TO output channels) register tiles; thread = (item, pixel group g of k), group g
// takes the tile columns g, g+k, ...  8 x 8 tiles: four LDS.128 feed 64 FMAs (the first version, 4 x 4 tiles, was bound by
// shared-memory wavefronts: 8 per 16 FMAs); the k groups keep all 288 threads busy although a 16->16 layer has only
// 36 items.  The groups' partial sums meet in shared memory before the one global atomic per weight and CTA.
template <int TI, int TO>
__global__ void __launch_bounds__(WS_MAXT, 2) conv3x3_wgrad_small_kernel(const WgradParams p, int tiles_x, int tiles_y, int n_tiles,
                                                                         int n_items, int k) {
    extern __shared__ __align__(16) float ws_smem[];
    const int Cgi = (p.Cin + TI - 1) / TI, Cgo = (p.Cout + TO - 1) / TO, Cip = Cgi * TI, Cop = Cgo * TO;
    const int C4i = Cip >> 2, C4o = Cop >> 2;
    const int prow = (WS_TH - 1) * p.stride + 3, pcol = (WS_TW - 1) * p.stride + 3;
    const int tid = threadIdx.x, nthr = blockDim.x;       // smem: 2 x ([prow][pcol][Cip] patch + [WS_PIX][Cop] dy tile)
    const int grp = tid / n_items, item = tid - grp * n_items;
    const bool live = grp < k;
    const int tap = item / (Cgi * Cgo), rem = item - tap * (Cgi * Cgo);
    const int cig = rem / Cgo, cog = rem - cig * Cgo;
    const int xoff = ((tap / 3) * pcol + (tap % 3)) * Cip + cig * TI, doff = cog * TO;
    float acc[TI][TO];
#pragma unroll
    for (int a = 0; a < TI; ++a)
#pragma unroll
        for (int b = 0; b < TO; ++b) acc[a][b] = 0.f;
    const int xstep = p.stride * Cip;
    const int bc = tid % Cop, bg = tid / Cop, nbg = nthr / Cop;   // bias gradient: thread = (channel, pixel group)
    float bsum = 0.f;
    // Two tile buffers filled with cp.async (zero fill outside the image): the next tile streams in while this one is
    // reduced -- with one buffer the kernel was bound by the load latency (0.5 TB/s on the level-1 layers).
    const int stage_floats = prow * pcol * Cip + WS_PIX * Cop;
    for (int e = tid; e < 2 * stage_floats; e += nthr) ws_smem[e] = 0.f;   // channel padding stays zero
    __syncthreads();
    const int s_c4 = tid % C4i, s_px = (tid / C4i) % pcol, s_py = tid / (C4i * pcol);
    const int d_c4 = nthr % C4i, d_px = (nthr / C4i) % pcol, d_py = nthr / (C4i * pcol);
    const int s_o4 = tid % C4o, s_pp = tid / C4o, d_o4 = nthr % C4o, d_pp = nthr / C4o;
    auto load_tile = [&](int t, float* patch, float* dys) {
        const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, b = t / (tiles_x * tiles_y);
        const int oy0 = ty * WS_TH, ox0 = tx * WS_TW;
        const int iy0 = oy0 * p.stride - p.pad_t, ix0 = ox0 * p.stride - p.pad_l;
        const float* xb = p.x + (size_t)b * p.H * p.W * p.x_cs;
        // (c4, px, py) of element e = tid + i * nthr advance incrementally: no divisions in the copy loops
        int c4 = s_c4, px = s_px, py = s_py;
        for (int e = tid; e < prow * pcol * C4i; e += nthr) {
            const int iy = iy0 + py, ix = ix0 + px;
            const bool ok = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
            const float* xp = ok ? xb + ((size_t)iy * p.W + ix) * p.x_cs + 4 * c4 : p.x;
            float* dst = patch + 4 * e;
            if (p.vec_x && 4 * c4 + 3 < p.Cin) cp_async16(dst, xp, ok);
            else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (4 * c4 + q < p.Cin) cp_async4(dst + q, ok ? xp + q : p.x, ok);
            }
            c4 += d_c4; px += d_px; py += d_py;
            if (c4 >= C4i) { c4 -= C4i; ++px; }
            if (px >= pcol) { px -= pcol; ++py; }
        }
        int o4 = s_o4, pp = s_pp;
        for (int e = tid; e < WS_PIX * C4o; e += nthr) {
            const int oy = oy0 + (pp >> 5), ox = ox0 + (pp & 31);
            const bool ok = oy < p.OH && ox < p.OW;
            const float* dp = ok ? p.dy + (((size_t)b * p.OH + oy) * p.OW + ox) * p.dy_cs + 4 * o4 : p.dy;
            float* dst = dys + 4 * e;
            if (p.vec_dy && 4 * o4 + 3 < p.Cout) cp_async16(dst, dp, ok);
            else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (4 * o4 + q < p.Cout) cp_async4(dst + q, ok ? dp + q : p.dy, ok);
            }
            o4 += d_o4; pp += d_pp;
            if (o4 >= C4o) { o4 -= C4o; ++pp; }
        }
        cp_async_commit();
    };
    int buf = 0;
    if ((int)blockIdx.x < n_tiles) load_tile(blockIdx.x, ws_smem, ws_smem + prow * pcol * Cip);
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, buf ^= 1) {
        const float* patch = ws_smem + buf * stage_floats;
        const float* dys = patch + prow * pcol * Cip;
        const int tn = t + gridDim.x;
        if (tn < n_tiles) {
            float* np = ws_smem + (buf ^ 1) * stage_floats;
            load_tile(tn, np, np + prow * pcol * Cip);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (live) {
#pragma unroll
            for (int r = 0; r < WS_TH; ++r) {
                const float* xr = patch + xoff + (r * p.stride * pcol) * Cip;
                const float* dr = dys + doff + (r * WS_TW) * Cop;
#pragma unroll 1
                for (int c = grp; c < WS_TW; c += k) {
                    float xa[TI], da[TO];
#pragma unroll
                    for (int q = 0; q < TI / 4; ++q) {
                        const float4 v = *reinterpret_cast<const float4*>(xr + c * xstep + 4 * q);
                        xa[4 * q] = v.x; xa[4 * q + 1] = v.y; xa[4 * q + 2] = v.z; xa[4 * q + 3] = v.w;
                    }
#pragma unroll
                    for (int q = 0; q < TO / 4; ++q) {
                        const float4 v = *reinterpret_cast<const float4*>(dr + c * Cop + 4 * q);
                        da[4 * q] = v.x; da[4 * q + 1] = v.y; da[4 * q + 2] = v.z; da[4 * q + 3] = v.w;
                    }
#pragma unroll
                    for (int a = 0; a < TI; ++a)
#pragma unroll
                        for (int b2 = 0; b2 < TO; ++b2) acc[a][b2] = fmaf(xa[a], da[b2], acc[a][b2]);
                }
            }
        }
        if (p.db && bg < nbg) {
            for (int pp = bg; pp < WS_PIX; pp += nbg) bsum += dys[pp * Cop + bc];
        }
        __syncthreads();                                 // this buffer is refilled by the next iteration's prefetch
    }
    // reduce the k pixel groups (and the bias groups) in shared memory, then one global atomic per weight
    __syncthreads();
    float* red = ws_smem;                                // [TI*TO][n_items] + [Cop]
    for (int e = tid; e < n_items * TI * TO + Cop; e += nthr) red[e] = 0.f;
    __syncthreads();
    if (live) {
#pragma unroll
        for (int a = 0; a < TI; ++a)
#pragma unroll
            for (int b2 = 0; b2 < TO; ++b2) atomicAdd(red + (a * TO + b2) * n_items + item, acc[a][b2]);   // item fastest: conflict-free
    }
    if (p.db && bg < nbg) atomicAdd(red + n_items * TI * TO + bc, bsum);
    __syncthreads();
    for (int e = tid; e < n_items * TI * TO; e += nthr) {
        const int ab = e / n_items, it = e - ab * n_items, a = ab / TO, b2 = ab - a * TO;
        const int tp = it / (Cgi * Cgo), rm = it - tp * (Cgi * Cgo);
        const int ci = (rm / Cgo) * TI + a, co = (rm % Cgo) * TO + b2;
        if (ci < p.Cin && co < p.Cout) atomicAdd(p.dw + ((size_t)tp * p.Cin + ci) * p.Cout + co, red[e]);
    }
    if (p.db && tid < p.Cout) atomicAdd(p.db + tid, red[n_items * TI * TO + tid]);
}

// ------------------------------------------------------------------------------------- element-wise helpers
__global__ void leaky_bwd_kernel(float* __restrict__ g, int g_cs, const float* __restrict__ y, int y_cs,
                                 size_t n_pix, int C, float alpha) {
    const size_t total = n_pix * C;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int c = idx % C; const size_t pix = idx / C;
        if (!(__ldg(y + pix * y_cs + c) > 0.f)) g[pix * g_cs + c] *= alpha;
    }
}

__global__ void add_strided_kernel(float* __restrict__ dst, int dst_cs, const float* __restrict__ src, int src_cs,
                                   size_t n_pix, int C, float scale) {
    const size_t total = n_pix * C;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int c = idx % C; const size_t pix = idx / C;
        dst[pix * dst_cs + c] += scale * __ldg(src + pix * src_cs + c);
    }
}

// ----------------------------------------------------------------------------------------- cost volume bwd
struct CvBwdParams {
    const float* g; const float* cv; const float* f0; const float* f1; const float* g_f0slot;
    float* df0; float* df1;
    int g_cs, cv_cs, f0_cs, f1_cs, gs_cs, df0_cs, df1_cs;
    int B, H, W, C, r;
    float alpha, inv_c;
    int acc_f1;
};

// One thread per (pixel, 4-channel group).  g' = g * leaky'(cv) is recomputed from the stored cost volume.
//   df0[p,c] += (1/C) sum_d g'[p,d] f1[p+d,c]  (+ the gradient that arrived through the f0 concat slot)
//   df1[q,c]  = (1/C) sum_d g'[q-d,d] f0[q-d,c]
__global__ void cost_volume_bwd_kernel(const CvBwdParams p) {
    const int C4 = p.C >> 2, nd1 = 2 * p.r + 1;
    const size_t total = (size_t)p.B * p.H * p.W * C4;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int k = idx % C4; const size_t pix = idx / C4;
        const int x = pix % p.W; const size_t row = pix / p.W;
        const int y = row % p.H; const size_t b = row / p.H;
        const size_t img = b * p.H * p.W;
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
        const float* gp = p.g + pix * p.g_cs;
        const float* cp = p.cv + pix * p.cv_cs;
        for (int v = -p.r; v <= p.r; ++v) {
            for (int h = -p.r; h <= p.r; ++h) {
                const int d = (v + p.r) * nd1 + (h + p.r);
                const int yy = y + v, xx = x + h;
                if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) {
                    float gv = __ldg(gp + d);
                    if (!(__ldg(cp + d) > 0.f)) gv *= p.alpha;
                    const float4 f = ldg4(p.f1 + (img + (size_t)yy * p.W + xx) * p.f1_cs + 4 * k);
                    a0.x = fmaf(gv, f.x, a0.x); a0.y = fmaf(gv, f.y, a0.y); a0.z = fmaf(gv, f.z, a0.z); a0.w = fmaf(gv, f.w, a0.w);
                }
                const int y2 = y - v, x2 = x - h;
                if (y2 >= 0 && y2 < p.H && x2 >= 0 && x2 < p.W) {
                    const size_t q = img + (size_t)y2 * p.W + x2;
                    float gv = __ldg(p.g + q * p.g_cs + d);
                    if (!(__ldg(p.cv + q * p.cv_cs + d) > 0.f)) gv *= p.alpha;
                    const float4 f = ldg4(p.f0 + q * p.f0_cs + 4 * k);
                    a1.x = fmaf(gv, f.x, a1.x); a1.y = fmaf(gv, f.y, a1.y); a1.z = fmaf(gv, f.z, a1.z); a1.w = fmaf(gv, f.w, a1.w);
                }
            }
        }
        float4* d0 = reinterpret_cast<float4*>(p.df0 + pix * p.df0_cs + 4 * k);
        float4 o = *d0;
        o.x += a0.x * p.inv_c; o.y += a0.y * p.inv_c; o.z += a0.z * p.inv_c; o.w += a0.w * p.inv_c;
        if (p.g_f0slot) {
            const float4 s = ldg4(p.g_f0slot + pix * p.gs_cs + 4 * k);
            o.x += s.x; o.y += s.y; o.z += s.z; o.w += s.w;
        }
        *d0 = o;
        float4* d1 = reinterpret_cast<float4*>(p.df1 + pix * p.df1_cs + 4 * k);
        float4 o1 = make_float4(a1.x * p.inv_c, a1.y * p.inv_c, a1.z * p.inv_c, a1.w * p.inv_c);
        if (p.acc_f1) { const float4 t = *d1; o1.x += t.x; o1.y += t.y; o1.z += t.z; o1.w += t.w; }
        *d1 = o1;
    }
}

// Tiled variant (search range 4).  The gather kernel above re-reads g, cv and the features through L1 for every (pixel,
// 4-channel group): 468 us at level 2 for 1 GFMA and 0.28 GB of algorithmic traffic.  Here one CTA owns an 8 x 16 pixel
// tile and a 32-channel slice, in one of two directions that share the arithmetic
//     out[p, c] = sum_d G[d][p] * F[p +- d][c]
//   DIR 0 (df0): G[d][p] = g'(p, d),      F = f1, offset +d
//   DIR 1 (df1): G[d][q] = g'(q - d, d),  F = f0, offset -d     (g' = g * leaky'(cv) / C, zero outside the image)
// G (81 x 128, pixel-fastest, the shift of DIR 1 applied while loading) and the 16 x 24 halo tile of F live in shared
// memory; a thread keeps a 4-pixel x 4-channel register tile and, per displacement row, a 12-pixel window of F, so 21
// LDS.128 feed 144 FMAs.
constexpr int CB_TH = 8, CB_TW = 16, CB_PIX = CB_TH * CB_TW, CB_CS = 32, CB_FS = CB_CS + 4, CB_FH = CB_TH + 8, CB_FW = CB_TW + 8,
              CB_GS = CB_PIX + 4, CB_THREADS = 256;
constexpr size_t CB_SMEM = ((size_t)81 * CB_GS + (size_t)CB_FH * CB_FW * CB_FS) * sizeof(float);

template <int DIR>
__device__ __forceinline__ void cv_bwd_tile(const CvBwdParams& p, float* __restrict__ Gs, float* __restrict__ Fs, int b, int y0, int x0,
                                            int c0) {
    const int tid = threadIdx.x;
    const size_t img = (size_t)b * p.H * p.W;
    // Copy loops: warp w owns tile row w; every lane issues a batch of independent loads (12-21 per array) before the
    // first shared-memory store, with no divisions by run-time values (the first versions were bound by load latency
    // and index arithmetic: 390 / 334 us at level 2).
    const int wid = tid >> 5, lane = tid & 31;
    const float inv_c = p.inv_c, inv_ca = p.inv_c * p.alpha;
    if (DIR == 0) {
        const int y = y0 + wid;
        const size_t rowpix = img + (size_t)y * p.W + x0;
#pragma unroll 1
        for (int p0 = 0; p0 < CB_TW; p0 += 4) {
            float gv[4][3], cvv[4][3];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool ok = y < p.H && x0 + p0 + i < p.W;
                const float* gp = p.g + (rowpix + p0 + i) * p.g_cs;
                const float* cp = p.cv + (rowpix + p0 + i) * p.cv_cs;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int d = lane + 32 * j;
                    gv[i][j] = 0.f; cvv[i][j] = 1.f;
                    if (ok && d < 81) { gv[i][j] = __ldg(gp + d); cvv[i][j] = __ldg(cp + d); }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int d = lane + 32 * j;
                    if (d < 81) Gs[d * CB_GS + wid * CB_TW + p0 + i] = gv[i][j] * (cvv[i][j] > 0.f ? inv_c : inv_ca);
                }
        }
    } else {
        // output row r = wid; for displacement row v the sources are row y0 + r - (v - 4), columns x0 - 4 .. x0 + 19,
        // group v of their 81 values (9 contiguous floats per pixel): element idx = c * 9 + h, 216 per (r, v)
#pragma unroll 1
        for (int v0 = 0; v0 < 9; v0 += 3) {
            float gv[3][7], cvv[3][7];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int v = v0 + i, sy = y0 + wid - (v - 4);
                const bool yok = sy >= 0 && sy < p.H;
                const size_t rowpix = img + (size_t)sy * p.W + (x0 - 4);
#pragma unroll
                for (int j = 0; j < 7; ++j) {
                    const int idx = lane + 32 * j, c = idx / 9, h = idx - c * 9, x = c - 8 + h, sx = x0 - 4 + c;
                    gv[i][j] = 0.f; cvv[i][j] = 1.f;
                    if (yok && idx < 216 && x >= 0 && x < CB_TW && sx >= 0 && sx < p.W) {
                        gv[i][j] = __ldg(p.g + (rowpix + c) * p.g_cs + v * 9 + h);
                        cvv[i][j] = __ldg(p.cv + (rowpix + c) * p.cv_cs + v * 9 + h);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 7; ++j) {
                    const int idx = lane + 32 * j, c = idx / 9, h = idx - c * 9, x = c - 8 + h;
                    if (idx < 216 && x >= 0 && x < CB_TW)
                        Gs[((v0 + i) * 9 + h) * CB_GS + wid * CB_TW + x] = gv[i][j] * (cvv[i][j] > 0.f ? inv_c : inv_ca);
                }
        }
    }
    {   // F halo tile: warp w copies rows 2w, 2w+1 (24 pixels x 8 float4 each)
        const float* F = DIR ? p.f0 : p.f1;
        const int f_cs = DIR ? p.f0_cs : p.f1_cs;
        const int c4 = lane & 7, ch = c0 + 4 * c4;
        float4 fv[2][6];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int fr = 2 * wid + i, y = y0 - 4 + fr;
            const bool yok = y >= 0 && y < p.H && ch < p.C;
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const int fc = (lane >> 3) + 4 * j, x = x0 - 4 + fc;
                fv[i][j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (yok && x >= 0 && x < p.W) fv[i][j] = ldg4(F + (img + (size_t)y * p.W + x) * f_cs + ch);
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 6; ++j)
                *reinterpret_cast<float4*>(Fs + ((2 * wid + i) * CB_FW + (lane >> 3) + 4 * j) * CB_FS + 4 * c4) = fv[i][j];
    }
    __syncthreads();

    const int cg = tid & 7, quad = tid >> 3, r = quad >> 2, xq = (quad & 3) * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 1
    for (int v = 0; v < 9; ++v) {
        const int frow = r + 4 + (DIR ? 4 - v : v - 4);
        const float* frp = Fs + (frow * CB_FW + xq) * CB_FS + cg * 4;
        float4 win[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) win[j] = *reinterpret_cast<const float4*>(frp + j * CB_FS);
        const float* gp = Gs + (v * 9) * CB_GS + r * CB_TW + xq;
#pragma unroll
        for (int h = 0; h < 9; ++h) {
            const float4 gq = *reinterpret_cast<const float4*>(gp + h * CB_GS);
            const float ga[4] = {gq.x, gq.y, gq.z, gq.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 f = win[i + 4 + (DIR ? 4 - h : h - 4)];
                acc[i][0] = fmaf(ga[i], f.x, acc[i][0]); acc[i][1] = fmaf(ga[i], f.y, acc[i][1]);
                acc[i][2] = fmaf(ga[i], f.z, acc[i][2]); acc[i][3] = fmaf(ga[i], f.w, acc[i][3]);
            }
        }
    }
    const int y = y0 + r, ch = c0 + 4 * cg;
    if (y >= p.H || ch >= p.C) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int x = x0 + xq + i;
        if (x >= p.W) continue;
        const size_t pix = img + (size_t)y * p.W + x;
        float4 o = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        if (DIR == 0) {
            float4* d0 = reinterpret_cast<float4*>(p.df0 + pix * p.df0_cs + ch);
            const float4 t = *d0;
            o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
            if (p.g_f0slot) {
                const float4 sl = ldg4(p.g_f0slot + pix * p.gs_cs + ch);
                o.x += sl.x; o.y += sl.y; o.z += sl.z; o.w += sl.w;
            }
            *d0 = o;
        } else {
            float4* d1 = reinterpret_cast<float4*>(p.df1 + pix * p.df1_cs + ch);
            if (p.acc_f1) { const float4 t = *d1; o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w; }
            *d1 = o;
        }
    }
}

__global__ void __launch_bounds__(CB_THREADS, 2) cost_volume_bwd_tiled_kernel(const CvBwdParams p, int tiles_x, int tiles_y) {
    extern __shared__ __align__(16) float cb_smem[];
    float* Gs = cb_smem;
    float* Fs = cb_smem + 81 * CB_GS;
    const int t = blockIdx.x, tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, b = t / (tiles_x * tiles_y);
    if (blockIdx.z) cv_bwd_tile<1>(p, Gs, Fs, b, ty * CB_TH, tx * CB_TW, blockIdx.y * CB_CS);
    else cv_bwd_tile<0>(p, Gs, Fs, b, ty * CB_TH, tx * CB_TW, blockIdx.y * CB_CS);
}

// ------------------------------------------------------------------------------------------------ zero insertion
// out (B,H,W,C) dense: out[b, 2y+oy, 2x+ox, :] = dy[b,y,x,:], zeros elsewhere.  With oy = 1 - pad_top (same for x) the
// stride-2 dgrad becomes the stride-1 dgrad of `out`, which runs on the tcgen05 conv kernel (3/4 of its MMAs multiply
// zeros, still several times faster than the CUDA-core parity-class kernel).
__global__ void dilate2_kernel(const float* __restrict__ dy, int dy_cs, float* __restrict__ out, int B, int OH, int OW,
                               int C4, int H, int W, int oy, int ox) {
    const size_t total = (size_t)B * H * W * C4;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int k = idx % C4; const size_t pix = idx / C4;
        const int x = pix % W; const size_t row = pix / W;
        const int y = row % H; const size_t b = row / H;
        const int sy = y - oy, sx = x - ox;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (sy >= 0 && sx >= 0 && !(sy & 1) && !(sx & 1) && (sy >> 1) < OH && (sx >> 1) < OW)
            v = ldg4(dy + ((b * OH + (sy >> 1)) * OW + (sx >> 1)) * dy_cs + 4 * k);
        reinterpret_cast<float4*>(out)[idx] = v;
    }
}

// ------------------------------------------------------------------------------------------------ warp bwd
// One warp per pixel, lanes over channels.  dx: scatter-add of the four weighted taps (clamped indices, weights
// from the UN-clamped fractional part, modules.py:107-137); dflow: only through the weights (floor / clip / cast
// have zero gradient), times flow_scale (model.py:109).  Nearest: one tap, no flow gradient.
template <bool NEAREST>
__global__ void warp_bwd_kernel(const float* __restrict__ x, int x_cs, const float* __restrict__ flow, int flow_cs,
                                float flow_scale, const float* __restrict__ g, int g_cs, float* __restrict__ dx, int dx_cs,
                                float* __restrict__ dflow, int dflow_cs, int B, int H, int W, int C) {
    const size_t total = (size_t)B * H * W;
    const int lane = threadIdx.x & 31;
    const size_t warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t pix = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5; pix < total; pix += warps) {
        const int px = pix % W; const size_t row = pix / W;
        const int py = row % H; const size_t b = row / H;
        const float fx = __ldg(flow + pix * flow_cs) * flow_scale, fy = __ldg(flow + pix * flow_cs + 1) * flow_scale;
        const size_t img = b * H * W;
        const float* gp = g + pix * g_cs;
        if (NEAREST) {
            const int ix = min(max(px + (int)fx, 0), W - 1);
            const int iy = min(max(py + (int)fy, 0), H - 1);
            float* d = dx + (img + (size_t)iy * W + ix) * dx_cs;
            for (int c = lane; c < C; c += 32) atomicAdd(d + c, __ldg(gp + c));
        } else {
            const float fx0 = floorf(fx), fy0 = floorf(fy);
            const float fx1 = fx0 + 1.f, fy1 = fy0 + 1.f;
            const float wl = (float)(W - 1), hl = (float)(H - 1);
            const int gy0 = (int)fminf(fmaxf((float)py + fy0, 0.f), hl);
            const int gy1 = (int)fminf(fmaxf((float)py + fy1, 0.f), hl);
            const int gx0 = (int)fminf(fmaxf((float)px + fx0, 0.f), wl);
            const int gx1 = (int)fminf(fmaxf((float)px + fx1, 0.f), wl);
            const float wy0 = fy1 - fy, wy1 = fy - fy0, wx0 = fx1 - fx, wx1 = fx - fx0;
            const size_t o00 = img + (size_t)gy0 * W + gx0, o01 = img + (size_t)gy0 * W + gx1;
            const size_t o10 = img + (size_t)gy1 * W + gx0, o11 = img + (size_t)gy1 * W + gx1;
            float dfx = 0.f, dfy = 0.f;
            for (int c = lane; c < C; c += 32) {
                const float gv = __ldg(gp + c);
                const float x00 = __ldg(x + o00 * x_cs + c), x01 = __ldg(x + o01 * x_cs + c);
                const float x10 = __ldg(x + o10 * x_cs + c), x11 = __ldg(x + o11 * x_cs + c);
                atomicAdd(dx + o00 * dx_cs + c, wy0 * wx0 * gv);
                atomicAdd(dx + o01 * dx_cs + c, wy0 * wx1 * gv);
                atomicAdd(dx + o10 * dx_cs + c, wy1 * wx0 * gv);
                atomicAdd(dx + o11 * dx_cs + c, wy1 * wx1 * gv);
                dfx = fmaf(gv, wy0 * (x01 - x00) + wy1 * (x11 - x10), dfx);
                dfy = fmaf(gv, wx0 * (x10 - x00) + wx1 * (x11 - x01), dfy);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                dfx += __shfl_xor_sync(0xffffffffu, dfx, o);
                dfy += __shfl_xor_sync(0xffffffffu, dfy, o);
            }
            if (lane == 0 && dflow) {
                dflow[pix * dflow_cs] += flow_scale * dfx;
                dflow[pix * dflow_cs + 1] += flow_scale * dfy;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- resize bwd
// Exact adjoint of resize_bilinear_kernel (warp_resize_loss.cu): each output scalar scatters its four weights.
__global__ void resize_bilinear_bwd_kernel(const float* __restrict__ g, int g_cs, float* __restrict__ dx, int dx_cs,
                                           int B, int H, int W, int C, int OH, int OW, float sy, float sx, float mul) {
    const size_t total = (size_t)B * OH * OW * C;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int c = idx % C;
        size_t pix = idx / C;
        const int ox = pix % OW; const size_t row = pix / OW;
        const int oy = row % OH; const size_t b = row / OH;
        const float fy = (float)oy * sy, fx = (float)ox * sx;
        const int ylo = (int)floorf(fy), xlo = (int)floorf(fx);
        const int yhi = min(ylo + 1, H - 1), xhi = min(xlo + 1, W - 1);
        const float yl = fy - (float)ylo, xl = fx - (float)xlo;
        const float gv = __ldg(g + pix * g_cs + c) * mul;
        float* db = dx + b * H * W * dx_cs + c;
        const float wtl = (1.f - xl) * (1.f - yl), wtr = xl * (1.f - yl), wbl = (1.f - xl) * yl, wbr = xl * yl;
        atomicAdd(db + ((size_t)ylo * W + xlo) * dx_cs, wtl * gv);
        if (wtr != 0.f) atomicAdd(db + ((size_t)ylo * W + xhi) * dx_cs, wtr * gv);
        if (wbl != 0.f) atomicAdd(db + ((size_t)yhi * W + xlo) * dx_cs, wbl * gv);
        if (wbr != 0.f) atomicAdd(db + ((size_t)yhi * W + xhi) * dx_cs, wbr * gv);
    }
}

// ------------------------------------------------------------------------------------------------ loss bwd
// d/dfs of weight/B * sum || gt_s - fs ||_ord : ord 2 -> (fs - gt_s)/||.||, ord 1 -> sign(fs - gt_s).
__global__ void lploss_level_bwd_kernel(const float* __restrict__ gt, int H, int W, const float* __restrict__ fs, int fs_cs,
                                        int h, int w, int B, float sy, float sx, float gt_div, float scale, int ord,
                                        float* __restrict__ gfs, int gfs_cs, int accumulate) {
    const size_t total = (size_t)B * h * w;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int x = idx % w; const size_t row = idx / w;
        const int y = row % h; const size_t b = row / h;
        const int iy = min((int)floorf((float)y * sy), H - 1), ix = min((int)floorf((float)x * sx), W - 1);
        const float2 gg = __ldg(reinterpret_cast<const float2*>(gt + ((b * H + iy) * W + ix) * 2));
        const float* f = fs + idx * fs_cs;
        const float dx = f[0] - gg.x / gt_div, dy = f[1] - gg.y / gt_div;
        float gx, gy;
        if (ord == 1) {
            gx = dx > 0.f ? scale : (dx < 0.f ? -scale : 0.f);
            gy = dy > 0.f ? scale : (dy < 0.f ? -scale : 0.f);
        } else {
            const float n = sqrtf(dx * dx + dy * dy);
            gx = scale * dx / n; gy = scale * dy / n;     // inf/NaN at exactly 0 like tf.norm's gradient
        }
        float* o = gfs + idx * gfs_cs;
        if (accumulate) { o[0] += gx; o[1] += gy; } else { o[0] = gx; o[1] = gy; }
    }
}

// ---------------------------------------------------------------------------------------------------- Adam
// hyper[0] = lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t) (computed by the host for step t, TF formula).
__global__ void adam_kernel(float* __restrict__ var, const float* __restrict__ grad, float* __restrict__ m,
                            float* __restrict__ v, size_t n, const float* __restrict__ hyper, float beta1, float beta2,
                            float eps, float gamma, float grad_scale) {
    const float lr_t = __ldg(hyper);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float w = var[i];
        const float g = grad[i] * grad_scale + gamma * w;
        const float mi = beta1 * m[i] + (1.f - beta1) * g;
        const float vi = beta2 * v[i] + (1.f - beta2) * g * g;
        m[i] = mi; v[i] = vi;
        var[i] = w - lr_t * mi / (sqrtf(vi) + eps);
    }
}

__global__ void sumsq_kernel(const float* __restrict__ x, size_t n, float scale, float* acc) {
    float s = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = __ldg(x + i);
        s = fmaf(v, v, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ float red[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) red[w] = s;
    __syncthreads();
    if (w == 0) {
        s = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) atomicAdd(acc, s * scale);
    }
}

__global__ void permute_cin_kernel(const float* __restrict__ src, float* __restrict__ dst, const int* __restrict__ perm,
                                   int cin_dst, int cin_src, int cout) {
    const size_t total = (size_t)9 * cin_dst * cout;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int co = idx % cout; const size_t r = idx / cout;
        const int ci = r % cin_dst; const int tap = r / cin_dst;
        const int s = __ldg(perm + ci);
        dst[idx] = s >= 0 ? __ldg(src + ((size_t)tap * cin_src + s) * cout + co) : 0.f;
    }
}

// w_rot[ky,kx,co,j] = w[2-ky,2-kx,ci_begin+j,co] for j < ci_count, 0 for ci_count <= j < ci_pad: the kernel with
// which a stride-1 SAME conv of dy gives dx[..., ci_begin : ci_begin+ci_count].
__global__ void rot_weights_kernel(const float* __restrict__ w, float* __restrict__ w_rot, int cin, int cout,
                                   int ci_begin, int ci_count, int ci_pad) {
    const size_t total = (size_t)9 * ci_pad * cout;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int j = idx % ci_pad; const size_t r = idx / ci_pad;
        const int co = r % cout; const int tap = r / cout;
        w_rot[idx] = j < ci_count ? __ldg(w + ((size_t)(8 - tap) * cin + ci_begin + j) * cout + co) : 0.f;
    }
}

}  // namespace pwc

using namespace pwc;

extern "C" int pwc_conv3x3_dgrad(const float* dy, int dy_cs, const float* w_hwio, float* dx, int dx_cs,
                                 const float* mask, int mask_cs, float mask_alpha, int accumulate,
                                 int B, int H, int W, int Cin, int Cout, int stride, int dilation, void* stream) {
    PWC_REQUIRE(dy && w_hwio && dx, PWC_E_BADARG, "conv3x3_dgrad: null pointer");
    PWC_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && stride >= 1 && dilation >= 1, PWC_E_BADARG,
                "conv3x3_dgrad: bad dims");
    PWC_REQUIRE(dy_cs >= Cout && dx_cs >= Cin && (!mask || mask_cs >= Cin), PWC_E_BADARG,
                "conv3x3_dgrad: channel stride smaller than channel count");
    PWC_REQUIRE(B <= 65535, PWC_E_BADARG, "conv3x3_dgrad: batch > 65535");
    DgradParams p{};
    p.dy = dy; p.w = w_hwio; p.mask = mask; p.dx = dx;
    p.dy_cs = dy_cs; p.mask_cs = mask_cs; p.dx_cs = dx_cs;
    p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.stride = stride; p.dil = dilation;
    same_pad_b(H, stride, dilation, &p.OH, &p.pad_t);
    same_pad_b(W, stride, dilation, &p.OW, &p.pad_l);
    p.mask_alpha = mask_alpha; p.accumulate = accumulate;
    p.vec_dy = aligned16(dy) && (dy_cs % 4 == 0);
    PWC_REQUIRE(stride <= 8, PWC_E_BADARG, "conv3x3_dgrad: stride > 8");
    const int Hs = (H + stride - 1) / stride, Ws = (W + stride - 1) / stride;     // largest parity class
    const int tiles = ((Ws + DG_TW - 1) / DG_TW) * ((Hs + DG_TH - 1) / DG_TH);
    dim3 grid(tiles, ((Cin + DG_BN - 1) / DG_BN) * stride * stride, B);
    conv3x3_dgrad_kernel<<<grid, DG_THREADS, 0, (cudaStream_t)stream>>>(p);
    PWC_CHECK_LAUNCH("conv3x3_dgrad_kernel");
    return 0;
}

extern "C" int pwc_conv3x3_wgrad(const float* x, int x_cs, const float* dy, int dy_cs, float* dw, float* db,
                                 const int* cin_map, int dw_cin, int B, int H, int W, int Cin, int Cout,
                                 int stride, int dilation, void* stream) {
    PWC_REQUIRE(x && dy && dw, PWC_E_BADARG, "conv3x3_wgrad: null pointer");
    PWC_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && stride >= 1 && dilation >= 1 && dw_cin > 0,
                PWC_E_BADARG, "conv3x3_wgrad: bad dims");
    PWC_REQUIRE(x_cs >= Cin && dy_cs >= Cout, PWC_E_BADARG, "conv3x3_wgrad: channel stride smaller than channel count");
    PWC_REQUIRE(cin_map || dw_cin == Cin, PWC_E_BADARG, "conv3x3_wgrad: dw_cin must equal Cin without a cin_map");
    WgradParams p{};
    p.x = x; p.dy = dy; p.dw = dw; p.db = db; p.cin_map = cin_map;
    p.x_cs = x_cs; p.dy_cs = dy_cs;
    p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.stride = stride; p.dil = dilation; p.dw_cin = dw_cin;
    same_pad_b(H, stride, dilation, &p.OH, &p.pad_t);
    same_pad_b(W, stride, dilation, &p.OW, &p.pad_l);
    p.total = (long long)B * p.OH * p.OW;
    p.vec_x = aligned16(x) && (x_cs % 4 == 0);
    p.vec_dy = aligned16(dy) && (dy_cs % 4 == 0);
    if (Cin <= 32 && Cout <= 32 && dilation == 1 && stride <= 2 && !cin_map) {
        const int tiles_x = (p.OW + WS_TW - 1) / WS_TW, tiles_y = (p.OH + WS_TH - 1) / WS_TH;
        const long long n_tiles = (long long)tiles_x * tiles_y * B;
        if (n_tiles < (1ll << 30)) {
            const int TI = Cin <= 4 ? 4 : 8, TO = Cout <= 4 ? 4 : 8;
            const int Cip = (Cin + TI - 1) / TI * TI, Cop = (Cout + TO - 1) / TO * TO;
            const int prow = (WS_TH - 1) * stride + 3, pcol = (WS_TW - 1) * stride + 3;
            const int n_items = 9 * (Cip / TI) * (Cop / TO);                 // <= 144
            size_t smem = 2 * ((size_t)prow * pcol * Cip + (size_t)WS_PIX * Cop) * 4;   // two tile buffers
            const size_t red = ((size_t)n_items * TI * TO + Cop) * 4;       // final reduction reuses the tile buffers
            if (smem < red) smem = red;
            int k = WS_MAXT / n_items;                                       // pixel groups
            if (k > WS_TW) k = WS_TW;
            const int threads = (n_items * k + 31) / 32 * 32;
            const int grid = (int)(n_tiles < 148 * 2 ? n_tiles : 148 * 2);
            auto launch = [&](auto kern) -> cudaError_t {
                cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return e;
                kern<<<grid, threads, smem, (cudaStream_t)stream>>>(p, tiles_x, tiles_y, (int)n_tiles, n_items, k);
                return cudaSuccess;
            };
            cudaError_t e = TI == 4 ? (TO == 4 ? launch(conv3x3_wgrad_small_kernel<4, 4>) : launch(conv3x3_wgrad_small_kernel<4, 8>))
                                    : (TO == 4 ? launch(conv3x3_wgrad_small_kernel<8, 4>) : launch(conv3x3_wgrad_small_kernel<8, 8>));
            if (e != cudaSuccess) { set_error("conv3x3_wgrad_small: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
            PWC_CHECK_LAUNCH("conv3x3_wgrad_small_kernel");
            return 0;
        }
    }
    // 128-wide tiles (8 registers per dimension) unless the 64-wide tiling pads noticeably less
    auto pick = [](int c) { const int p128 = (c + 127) / 128 * 128, p64 = (c + 63) / 64 * 64; return p128 * 100 <= p64 * 115 ? 128 : 64; };
    const int TM = pick(Cin), TN = pick(Cout);
    p.ci_tiles = (Cin + TM - 1) / TM;
    const int co_tiles = (Cout + TN - 1) / TN;
    const int tiles = p.ci_tiles * co_tiles;
    PWC_REQUIRE(tiles <= 65535, PWC_E_BADARG, "conv3x3_wgrad: too many channel tiles");
    // ~6 CTAs per SM in total; every CTA reduces at least 64 pixels
    long long chunks = (148LL * 6 + tiles * 9 - 1) / (tiles * 9);
    const long long max_chunks = (p.total + 63) / 64;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    p.chunk = ((p.total + chunks - 1) / chunks + 15) / 16 * 16;
    chunks = (p.total + p.chunk - 1) / p.chunk;
    dim3 grid((unsigned)chunks, tiles, 9);
    cudaStream_t st = (cudaStream_t)stream;
    if (TM == 128 && TN == 128) conv3x3_wgrad_kernel<128, 128><<<grid, WG_THREADS, 0, st>>>(p);
    else if (TM == 128) conv3x3_wgrad_kernel<128, 64><<<grid, WG_THREADS, 0, st>>>(p);
    else if (TN == 128) conv3x3_wgrad_kernel<64, 128><<<grid, WG_THREADS, 0, st>>>(p);
    else conv3x3_wgrad_kernel<64, 64><<<grid, WG_THREADS, 0, st>>>(p);
    PWC_CHECK_LAUNCH("conv3x3_wgrad_kernel");
    return 0;
}

extern "C" int pwc_leaky_bwd(float* g, int g_cs, const float* y, int y_cs, long long n_pix, int C, float alpha, void* stream) {
    PWC_REQUIRE(g && y, PWC_E_BADARG, "leaky_bwd: null pointer");
    PWC_REQUIRE(n_pix > 0 && C > 0 && g_cs >= C && y_cs >= C, PWC_E_BADARG, "leaky_bwd: bad dims");
    leaky_bwd_kernel<<<grid_for((size_t)n_pix * C, 256), 256, 0, (cudaStream_t)stream>>>(g, g_cs, y, y_cs, (size_t)n_pix, C, alpha);
    PWC_CHECK_LAUNCH("leaky_bwd_kernel");
    return 0;
}

extern "C" int pwc_add_strided(float* dst, int dst_cs, const float* src, int src_cs, long long n_pix, int C, float scale,
                               void* stream) {
    PWC_REQUIRE(dst && src, PWC_E_BADARG, "add_strided: null pointer");
    PWC_REQUIRE(n_pix > 0 && C > 0 && dst_cs >= C && src_cs >= C, PWC_E_BADARG, "add_strided: bad dims");
    add_strided_kernel<<<grid_for((size_t)n_pix * C, 256), 256, 0, (cudaStream_t)stream>>>(dst, dst_cs, src, src_cs, (size_t)n_pix, C, scale);
    PWC_CHECK_LAUNCH("add_strided_kernel");
    return 0;
}

extern "C" int pwc_dilate2(const float* dy, int dy_cs, float* out, int B, int OH, int OW, int C, int H, int W, int oy, int ox,
                           void* stream) {
    PWC_REQUIRE(dy && out, PWC_E_BADARG, "dilate2: null pointer");
    PWC_REQUIRE(B > 0 && OH > 0 && OW > 0 && C > 0 && H > 0 && W > 0 && dy_cs >= C, PWC_E_BADARG, "dilate2: bad dims");
    PWC_REQUIRE((C & 3) == 0 && (dy_cs & 3) == 0 && aligned16(dy) && aligned16(out), PWC_E_ALIGN,
                "dilate2: C and the channel stride must be multiples of 4, pointers 16-byte aligned");
    dilate2_kernel<<<grid_for((size_t)B * H * W * (C >> 2), 256), 256, 0, (cudaStream_t)stream>>>(dy, dy_cs, out, B, OH, OW, C >> 2,
                                                                                                 H, W, oy, ox);
    PWC_CHECK_LAUNCH("dilate2_kernel");
    return 0;
}

extern "C" int pwc_cost_volume_bwd(const float* g, int g_cs, const float* cv, int cv_cs, const float* f0, int f0_cs,
                                   const float* f1, int f1_cs, const float* g_f0slot, int gs_cs,
                                   float* df0, int df0_cs, float* df1, int df1_cs, int accumulate_f1,
                                   int B, int H, int W, int C, int search_range, float alpha, void* stream) {
    PWC_REQUIRE(g && cv && f0 && f1 && df0 && df1, PWC_E_BADARG, "cost_volume_bwd: null pointer");
    PWC_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && search_range >= 0, PWC_E_BADARG, "cost_volume_bwd: bad dims");
    const int nd = (2 * search_range + 1) * (2 * search_range + 1);
    PWC_REQUIRE(g_cs >= nd && cv_cs >= nd, PWC_E_BADARG, "cost_volume_bwd: channel stride smaller than (2r+1)^2");
    PWC_REQUIRE((C & 3) == 0 && (f0_cs & 3) == 0 && (f1_cs & 3) == 0 && (df0_cs & 3) == 0 && (df1_cs & 3) == 0 &&
                aligned16(f0) && aligned16(f1) && aligned16(df0) && aligned16(df1) &&
                (!g_f0slot || (aligned16(g_f0slot) && (gs_cs & 3) == 0)), PWC_E_ALIGN,
                "cost_volume_bwd: C and feature strides must be multiples of 4, feature pointers 16-byte aligned");
    CvBwdParams p{};
    p.g = g; p.cv = cv; p.f0 = f0; p.f1 = f1; p.g_f0slot = g_f0slot; p.df0 = df0; p.df1 = df1;
    p.g_cs = g_cs; p.cv_cs = cv_cs; p.f0_cs = f0_cs; p.f1_cs = f1_cs; p.gs_cs = gs_cs; p.df0_cs = df0_cs; p.df1_cs = df1_cs;
    p.B = B; p.H = H; p.W = W; p.C = C; p.r = search_range; p.alpha = alpha; p.inv_c = 1.f / (float)C;
    p.acc_f1 = accumulate_f1;
    static const bool gather = [] { const char* e = getenv("PWC_CV_BWD"); return e && !strcmp(e, "gather"); }();
    const long long tiles_x = (W + CB_TW - 1) / CB_TW, tiles_y = (H + CB_TH - 1) / CB_TH;
    if (search_range == 4 && !gather && tiles_x * tiles_y * B < (1ll << 31) && (C + CB_CS - 1) / CB_CS <= 65535) {
        cudaError_t e = cudaFuncSetAttribute(cost_volume_bwd_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CB_SMEM);
        if (e != cudaSuccess) { set_error("cost_volume_bwd: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        dim3 grid((unsigned)(tiles_x * tiles_y * B), (C + CB_CS - 1) / CB_CS, 2);
        cost_volume_bwd_tiled_kernel<<<grid, CB_THREADS, CB_SMEM, (cudaStream_t)stream>>>(p, (int)tiles_x, (int)tiles_y);
        PWC_CHECK_LAUNCH("cost_volume_bwd_tiled_kernel");
        return 0;
    }
    const size_t total = (size_t)B * H * W * (C / 4);
    cost_volume_bwd_kernel<<<grid_for(total, 128, 148 * 32), 128, 0, (cudaStream_t)stream>>>(p);
    PWC_CHECK_LAUNCH("cost_volume_bwd_kernel");
    return 0;
}

extern "C" int pwc_warp_bwd(const float* x, int x_cs, const float* flow, int flow_cs, float flow_scale, int warp_type,
                            const float* g, int g_cs, float* dx, int dx_cs, float* dflow, int dflow_cs,
                            int B, int H, int W, int C, void* stream) {
    PWC_REQUIRE(x && flow && g && dx, PWC_E_BADARG, "warp_bwd: null pointer");
    PWC_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, PWC_E_BADARG, "warp_bwd: bad dims");
    PWC_REQUIRE(warp_type == 0 || warp_type == 1, PWC_E_BADARG, "warp_bwd: warp_type must be 0 (bilinear) or 1 (nearest)");
    const size_t total = (size_t)B * H * W * 32;
    const int blocks = grid_for(total, 256);
    if (warp_type == 1)
        warp_bwd_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, x_cs, flow, flow_cs, flow_scale, g, g_cs, dx, dx_cs, dflow, dflow_cs, B, H, W, C);
    else
        warp_bwd_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, x_cs, flow, flow_cs, flow_scale, g, g_cs, dx, dx_cs, dflow, dflow_cs, B, H, W, C);
    PWC_CHECK_LAUNCH("warp_bwd_kernel");
    return 0;
}

extern "C" int pwc_resize_bilinear_bwd(const float* g, int g_cs, float* dx, int dx_cs, int B, int H, int W, int C,
                                       int OH, int OW, float mul, void* stream) {
    PWC_REQUIRE(g && dx, PWC_E_BADARG, "resize_bwd: null pointer");
    PWC_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && OH > 0 && OW > 0, PWC_E_BADARG, "resize_bwd: bad dims");
    const size_t total = (size_t)B * OH * OW * C;
    const float sy = (float)H / (float)OH, sx = (float)W / (float)OW;
    resize_bilinear_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(g, g_cs, dx, dx_cs, B, H, W, C, OH, OW, sy, sx, mul);
    PWC_CHECK_LAUNCH("resize_bilinear_bwd_kernel");
    return 0;
}

extern "C" int pwc_lploss_level_bwd(const float* gt, int H, int W, const float* fs, int fs_cs, int h, int w, int B,
                                    float gt_div, float weight, int ord, float* gfs, int gfs_cs, int accumulate,
                                    void* stream) {
    PWC_REQUIRE(gt && fs && gfs, PWC_E_BADARG, "lploss_bwd: null pointer");
    PWC_REQUIRE(ord == 1 || ord == 2, PWC_E_BADARG, "lploss_bwd: ord must be 1 or 2");
    PWC_REQUIRE(B > 0 && H > 0 && W > 0 && h > 0 && w > 0 && fs_cs >= 2 && gfs_cs >= 2, PWC_E_BADARG, "lploss_bwd: bad dims");
    const size_t total = (size_t)B * h * w;
    lploss_level_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        gt, H, W, fs, fs_cs, h, w, B, (float)H / (float)h, (float)W / (float)w, gt_div, weight / (float)B, ord, gfs, gfs_cs,
        accumulate);
    PWC_CHECK_LAUNCH("lploss_level_bwd_kernel");
    return 0;
}

extern "C" int pwc_adam_step(float* var, const float* grad, float* m, float* v, long long n, const float* lr_t_dev,
                             float beta1, float beta2, float eps, float gamma, float grad_scale, void* stream) {
    PWC_REQUIRE(var && grad && m && v && lr_t_dev, PWC_E_BADARG, "adam_step: null pointer");
    PWC_REQUIRE(n > 0, PWC_E_BADARG, "adam_step: n must be positive");
    adam_kernel<<<grid_for((size_t)n, 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(var, grad, m, v, (size_t)n, lr_t_dev, beta1,
                                                                                  beta2, eps, gamma, grad_scale);
    PWC_CHECK_LAUNCH("adam_kernel");
    return 0;
}

extern "C" int pwc_sumsq(const float* x, long long n, float scale, float* acc, void* stream) {
    PWC_REQUIRE(x && acc && n > 0, PWC_E_BADARG, "sumsq: bad arguments");
    sumsq_kernel<<<grid_for((size_t)n, 256, 148 * 4), 256, 0, (cudaStream_t)stream>>>(x, (size_t)n, scale, acc);
    PWC_CHECK_LAUNCH("sumsq_kernel");
    return 0;
}

extern "C" int pwc_permute_cin(const float* w_src, float* w_dst, const int* perm, int cin_dst, int cin_src, int cout,
                               void* stream) {
    PWC_REQUIRE(w_src && w_dst && perm && cin_dst > 0 && cin_src > 0 && cout > 0, PWC_E_BADARG, "permute_cin: bad arguments");
    permute_cin_kernel<<<grid_for((size_t)9 * cin_dst * cout, 256), 256, 0, (cudaStream_t)stream>>>(w_src, w_dst, perm, cin_dst, cin_src, cout);
    PWC_CHECK_LAUNCH("permute_cin_kernel");
    return 0;
}

extern "C" int pwc_conv3x3_rot_weights(const float* w_hwio, float* w_rot, int Cin, int Cout, int ci_begin, int ci_count,
                                       int ci_pad, void* stream) {
    PWC_REQUIRE(w_hwio && w_rot && Cin > 0 && Cout > 0, PWC_E_BADARG, "rot_weights: bad arguments");
    PWC_REQUIRE(ci_begin >= 0 && ci_count > 0 && ci_begin + ci_count <= Cin && ci_pad >= ci_count, PWC_E_BADARG,
                "rot_weights: bad channel range");
    rot_weights_kernel<<<grid_for((size_t)9 * ci_pad * Cout, 256), 256, 0, (cudaStream_t)stream>>>(w_hwio, w_rot, Cin, Cout,
                                                                                                 ci_begin, ci_count, ci_pad);
    PWC_CHECK_LAUNCH("rot_weights_kernel");
    return 0;
}
