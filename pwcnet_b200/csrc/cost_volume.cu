// Cost volume (+ optional fused feature warp) for PWC-Net, sm_100a.
//
// Replaces CostVolumeLayer.__call__ / get_cost (reference modules.py:164-204) -- 81 x
// (pad, pad, mul, crop, mean) graph nodes per level -- and, in the fused entry point, also
// WarpingLayer / bilinear_warp / nearest_warp (modules.py:83-154) by ONE kernel per level:
//
//   out[b,y,x,(v+4)*9+(h+4)] = leaky( (1/C) * sum_c f0[b,y,x,c] * f1w[b,y+v,x+h,c] )
//
// HBM-bound by design: f0 and f1 are read once (the f1 halo re-reads hit L2), the 81-channel
// result is written once, straight into the estimator's concat buffer.
//
// Tiling (R = 4): a CTA (128 threads) owns a TY x TX = 6 x 32 output tile.  Per 32-channel chunk it
// stages the f0 tile and the (TY+8) x (TX+8) f1 halo tile in shared memory (cp.async with zero fill
// outside the image = the reference's zero padding; in the fused variant the halo pixels are
// bilinearly gathered from f1 on the fly).  Work decomposition: all (row y, vertical shift v) pairs
// that read the SAME f1 row r = y + v are grouped, and a thread owns one f1 row r, two consecutive
// output rows (y, y+1) of that group and an 8-pixel strip: 2 x 8 x 9 = 144 accumulators.  Per 4
// channels it issues 16 (f1 row, shared by both output rows) + 2 x 8 (f0) LDS.128 for 576 FMAs,
// i.e. 4.5 FMAs per shared-memory word: above the 4.0 needed to be FMA-bound rather than LDS-bound
// (an LDS.128 costs four shared-memory wavefronts whatever the lane addresses are).  The 30 work items
// of a strip fill one warp; lanes of a quarter-warp touch distinct bank groups (rows padded by 16 B).
// Results are transposed through shared memory so the 81 channels of a pixel leave as one contiguous
// 324-byte run.
#include "cost_volume.cuh"
#include <cstdlib>

namespace pwc {

constexpr int CV_R = 4;
constexpr int CV_D = 2 * CV_R + 1;     // 9
constexpr int CV_ND = CV_D * CV_D;     // 81
static_assert(CV_ND == 81, "r=4 kernel");
constexpr int CV_TY = 6;
#ifndef PWC_CV_TX
#define PWC_CV_TX 32
#endif
constexpr int CV_TX = PWC_CV_TX;
constexpr int CV_SX = 8;               // strip width (pixels per thread)
constexpr int CV_CH = 32;              // channels per chunk
constexpr int CV_HY = CV_TY + 2 * CV_R;   // 14
constexpr int CV_HX = CV_TX + 2 * CV_R;   // 40
constexpr int CV_F0_ROW = CV_TX * CV_CH + 4;   // floats; +4 -> rows land in distinct 16B bank groups
constexpr int CV_F1_ROW = CV_HX * CV_CH + 4;
constexpr int CV_F0_FLOATS = (CV_TY + 1) * CV_F0_ROW;   // +1 row: single items read a dummy second row
constexpr int CV_F1_FLOATS = CV_HY * CV_F1_ROW;
constexpr int CV_OPIX = 84;                    // smem floats per output pixel (81 padded to 16B multiple)
constexpr int CV_OROW = CV_TX * CV_OPIX + 8;   // + 8 floats: de-phase rows across banks
constexpr int CV_THREADS = 32 * (CV_TX / CV_SX);   // one warp per 8-pixel strip
constexpr int CV_CTAS_PER_SM = CV_TX == 32 ? 2 : 4;
constexpr int CV_ITEMS = 30;                   // (f1 row, output row pair) items per strip for TY = 6
constexpr int CV_SMEM_BYTES = (CV_F0_FLOATS + CV_F1_FLOATS) * 4;
constexpr int CV_TAPS_WORDS = CV_HY * CV_HX * 6;   // fused variants: per halo pixel 4 tap offsets + 2 fractions
constexpr int CV_SMEM_BYTES_FUSED = CV_SMEM_BYTES + CV_TAPS_WORDS * 4;
static_assert(CV_TY * CV_OROW <= CV_F1_FLOATS, "output staging must fit in the f1 halo buffer");


// Bilinear / nearest sample of 4 channels of f1 at pixel (y,x) displaced by the flow there.
// Index clamping and weights follow modules.py:107-137 (clamped taps, un-clamped weights).
template <int WARP>
__device__ __forceinline__ float4 sample_f1(const CvParams& p, const float* f1b, const float* flowb,
                                            int y, int x, int c) {
    if (WARP == 0) return ldg4(f1b + ((size_t)y * p.W + x) * p.f1_cs + c);
    const float* fl = flowb + ((size_t)y * p.W + x) * p.flow_cs;
    const float fx = __ldg(fl) * p.flow_scale, fy = __ldg(fl + 1) * p.flow_scale;
    if (WARP == 2) {  // nearest: tf.cast(flow, int32) truncates toward zero (modules.py:85)
        int ix = min(max(x + (int)fx, 0), p.W - 1);
        int iy = min(max(y + (int)fy, 0), p.H - 1);
        return ldg4(f1b + ((size_t)iy * p.W + ix) * p.f1_cs + c);
    }
    const float fx0 = floorf(fx), fy0 = floorf(fy);
    const float fx1 = fx0 + 1.f, fy1 = fy0 + 1.f;
    const float wl = (float)(p.W - 1), hl = (float)(p.H - 1);
    const int gy0 = (int)fminf(fmaxf((float)y + fy0, 0.f), hl);
    const int gy1 = (int)fminf(fmaxf((float)y + fy1, 0.f), hl);
    const int gx0 = (int)fminf(fmaxf((float)x + fx0, 0.f), wl);
    const int gx1 = (int)fminf(fmaxf((float)x + fx1, 0.f), wl);
    const float c00 = (fy1 - fy) * (fx1 - fx), c01 = (fy1 - fy) * (fx - fx0);
    const float c10 = (fy - fy0) * (fx1 - fx), c11 = (fy - fy0) * (fx - fx0);
    const float4 a = ldg4(f1b + ((size_t)gy0 * p.W + gx0) * p.f1_cs + c);
    const float4 b = ldg4(f1b + ((size_t)gy0 * p.W + gx1) * p.f1_cs + c);
    const float4 d = ldg4(f1b + ((size_t)gy1 * p.W + gx0) * p.f1_cs + c);
    const float4 e = ldg4(f1b + ((size_t)gy1 * p.W + gx1) * p.f1_cs + c);
    float4 r;
    r.x = c00 * a.x + c01 * b.x + c10 * d.x + c11 * e.x;
    r.y = c00 * a.y + c01 * b.y + c10 * d.y + c11 * e.y;
    r.z = c00 * a.z + c01 * b.z + c10 * d.z + c11 * e.z;
    r.w = c00 * a.w + c01 * b.w + c10 * d.w + c11 * e.w;
    return r;
}

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;   // src-size 0 -> 16 bytes of zeros
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}

// item -> (f1 halo row index 0..13, first output row ya, rows n in {1,2}); items sorted by f1 row
__device__ __forceinline__ void cv_item(int item, int& r_idx, int& ya, int& n) {
    int cnt = 0;
    r_idx = 0; ya = 0; n = 0;
#pragma unroll 1
    for (int r = 0; r < CV_HY; ++r) {
        const int ylo = max(0, r - 2 * CV_R), yhi = min(CV_TY - 1, r);   // |(r - 4) - y| <= 4
        const int rows = yhi - ylo + 1, items = (rows + 1) >> 1;
        if (item < cnt + items) {
            const int j = item - cnt;
            r_idx = r; ya = ylo + 2 * j; n = min(2, yhi - ya + 1);
            return;
        }
        cnt += items;
    }
}

// WARP: 0 = f1 used as is, 1 = bilinear, 2 = nearest
template <int WARP>
__global__ void __launch_bounds__(CV_THREADS, CV_CTAS_PER_SM) cost_volume_r4_kernel(const CvParams p) {
    extern __shared__ __align__(16) float smem[];
    float* f0s = smem;
    float* f1s = smem + CV_F0_FLOATS;

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * CV_TX, y0 = blockIdx.y * CV_TY, b = blockIdx.z;
    const float* f0b = p.f0 + (size_t)b * p.H * p.W * p.f0_cs;
    const float* f1b = p.f1 + (size_t)b * p.H * p.W * p.f1_cs;
    const float* flowb = WARP ? p.flow + (size_t)b * p.H * p.W * p.flow_cs : nullptr;

    // work item of this thread: strip = warp, item = lane
    const int xs = tid >> 5, lane = tid & 31;
    int r_idx, ya, n_rows;
    cv_item(lane < CV_ITEMS ? lane : 0, r_idx, ya, n_rows);
    if (lane >= CV_ITEMS) n_rows = 0;

    float acc0[CV_SX][CV_D], acc1[CV_SX][CV_D];
#pragma unroll
    for (int i = 0; i < CV_SX; ++i)
#pragma unroll
        for (int j = 0; j < CV_D; ++j) { acc0[i][j] = 0.f; acc1[i][j] = 0.f; }

    // Fused variants: per halo pixel, resolve the flow once into 4 clamped tap offsets and the two
    // fractional weights (modules.py:107-123); the per-chunk gather below then has a single level of
    // dependent loads.  oy0 < 0 marks a halo pixel outside the image (zero padding of the cost volume).
    int* taps = reinterpret_cast<int*>(smem + CV_F0_FLOATS + CV_F1_FLOATS);
    if (WARP != 0) {
        for (int pix = tid; pix < CV_HY * CV_HX; pix += CV_THREADS) {
            const int py = pix / CV_HX, px = pix - py * CV_HX;
            const int gy = y0 + py - CV_R, gx = x0 + px - CV_R;
            int oy0 = -1, oy1 = 0, ox0 = 0, ox1 = 0;
            float ty = 0.f, tx = 0.f;
            if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) {
                const float* fl = flowb + ((size_t)gy * p.W + gx) * p.flow_cs;
                const float fx = __ldg(fl) * p.flow_scale, fy = __ldg(fl + 1) * p.flow_scale;
                int iy0, iy1, ix0, ix1;
                if (WARP == 2) {   // nearest: tf.cast(flow, int32) truncates toward zero (modules.py:85)
                    iy0 = iy1 = min(max(gy + (int)fy, 0), p.H - 1);
                    ix0 = ix1 = min(max(gx + (int)fx, 0), p.W - 1);
                } else {
                    const float fx0 = floorf(fx), fy0 = floorf(fy);
                    const float wl = (float)(p.W - 1), hl = (float)(p.H - 1);
                    iy0 = (int)fminf(fmaxf((float)gy + fy0, 0.f), hl);
                    iy1 = (int)fminf(fmaxf((float)gy + (fy0 + 1.f), 0.f), hl);
                    ix0 = (int)fminf(fmaxf((float)gx + fx0, 0.f), wl);
                    ix1 = (int)fminf(fmaxf((float)gx + (fx0 + 1.f), 0.f), wl);
                    ty = fy - fy0; tx = fx - fx0;   // exact; (f0 + 1) - f == 1 - t after one rounding
                }
                oy0 = iy0 * p.W * p.f1_cs; oy1 = iy1 * p.W * p.f1_cs;
                ox0 = ix0 * p.f1_cs; ox1 = ix1 * p.f1_cs;
            }
            int* t = taps + pix * 6;
            t[0] = oy0; t[1] = oy1; t[2] = ox0; t[3] = ox1;
            t[4] = __float_as_int(ty); t[5] = __float_as_int(tx);
        }
        __syncthreads();
    }

    for (int c0 = 0; c0 < p.C; c0 += CV_CH) {
        const int nch4 = min(CV_CH, p.C - c0) >> 2;   // float4 groups valid in this chunk
        if (c0) __syncthreads();
        // ---- stage the f0 tile
        for (int e = tid; e < CV_TY * CV_TX * (CV_CH / 4); e += CV_THREADS) {
            const int k = e & 7, px = (e >> 3) & (CV_TX - 1), py = e / (CV_TX * 8);
            const int gy = y0 + py, gx = x0 + px;
            const bool ok = gy < p.H && gx < p.W && k < nch4;
            const float* src = ok ? f0b + ((size_t)gy * p.W + gx) * p.f0_cs + c0 + 4 * k : f0b;
            cp_async16(f0s + py * CV_F0_ROW + px * CV_CH + 4 * k, src, ok);
        }
        // ---- stage the f1 halo tile (warped on the fly in the fused variants), zeros outside the image
        if (WARP == 0) {
            for (int e = tid; e < CV_HY * CV_HX * (CV_CH / 4); e += CV_THREADS) {
                const int k = e & 7, pix = e >> 3;
                const int py = pix / CV_HX, px = pix - py * CV_HX;
                const int gy = y0 + py - CV_R, gx = x0 + px - CV_R;
                const bool ok = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W && k < nch4;
                const float* src = ok ? f1b + ((size_t)gy * p.W + gx) * p.f1_cs + c0 + 4 * k : f1b;
                cp_async16(f1s + py * CV_F1_ROW + px * CV_CH + 4 * k, src, ok);
            }
        } else {
            const float* f1c = f1b + c0;
#pragma unroll 2
            for (int e = tid; e < CV_HY * CV_HX * (CV_CH / 4); e += CV_THREADS) {
                const int k = e & 7, pix = e >> 3;
                const int py = pix / CV_HX, px = pix - py * CV_HX;
                const int* t = taps + pix * 6;
                const int oy0 = t[0];
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (oy0 >= 0 && k < nch4) {
                    const int oy1 = t[1], ox0 = t[2], ox1 = t[3];
                    const float ty = __int_as_float(t[4]), tx = __int_as_float(t[5]);
                    const float4 a = ldg4(f1c + oy0 + ox0 + 4 * k), bq = ldg4(f1c + oy0 + ox1 + 4 * k);
                    const float4 d = ldg4(f1c + oy1 + ox0 + 4 * k), g = ldg4(f1c + oy1 + ox1 + 4 * k);
                    const float sy = 1.f - ty, sx = 1.f - tx;
                    const float c00 = sy * sx, c01 = sy * tx, c10 = ty * sx, c11 = ty * tx;
                    v.x = c00 * a.x + c01 * bq.x + c10 * d.x + c11 * g.x;
                    v.y = c00 * a.y + c01 * bq.y + c10 * d.y + c11 * g.y;
                    v.z = c00 * a.z + c01 * bq.z + c10 * d.z + c11 * g.z;
                    v.w = c00 * a.w + c01 * bq.w + c10 * d.w + c11 * g.w;
                }
                *reinterpret_cast<float4*>(f1s + py * CV_F1_ROW + px * CV_CH + 4 * k) = v;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        // ---- optional copy of the staged f0 tile into the estimator's concat slot
        if (p.f0_copy) {
            for (int e = tid; e < CV_TY * CV_TX * (CV_CH / 4); e += CV_THREADS) {
                const int k = e & 7, px = (e >> 3) & (CV_TX - 1), py = e / (CV_TX * 8);
                const int gy = y0 + py, gx = x0 + px;
                if (gy < p.H && gx < p.W && k < nch4)
                    *reinterpret_cast<float4*>(p.f0_copy + (((size_t)b * p.H + gy) * p.W + gx) * p.f0_copy_cs + c0 + 4 * k) =
                        *reinterpret_cast<const float4*>(f0s + py * CV_F0_ROW + px * CV_CH + 4 * k);
            }
        }
        // ---- correlate: f1 row r against output rows ya (v = r - ya) and ya + 1 (v - 1)
        const float* a0_base = f0s + ya * CV_F0_ROW + xs * CV_SX * CV_CH;
        const float* a1_base = a0_base + CV_F0_ROW;
        const float* f_base = f1s + r_idx * CV_F1_ROW + xs * CV_SX * CV_CH;
#pragma unroll 1
        for (int k = 0; k < CV_CH / 4; ++k) {
            float4 f[CV_D];   // sliding window over the f1 row: columns i .. i+8
#pragma unroll
            for (int q = 0; q < CV_D; ++q) f[q] = *reinterpret_cast<const float4*>(f_base + q * CV_CH + 4 * k);
#pragma unroll
            for (int i = 0; i < CV_SX; ++i) {
                const float4 a0 = *reinterpret_cast<const float4*>(a0_base + i * CV_CH + 4 * k);
                const float4 a1 = *reinterpret_cast<const float4*>(a1_base + i * CV_CH + 4 * k);
                // component-major order: 18 independent FMAs between two updates of the same accumulator
#pragma unroll
                for (int j = 0; j < CV_D; ++j) { const float fx = f[(i + j) % CV_D].x; acc0[i][j] = fmaf(a0.x, fx, acc0[i][j]); acc1[i][j] = fmaf(a1.x, fx, acc1[i][j]); }
#pragma unroll
                for (int j = 0; j < CV_D; ++j) { const float fy = f[(i + j) % CV_D].y; acc0[i][j] = fmaf(a0.y, fy, acc0[i][j]); acc1[i][j] = fmaf(a1.y, fy, acc1[i][j]); }
#pragma unroll
                for (int j = 0; j < CV_D; ++j) { const float fz = f[(i + j) % CV_D].z; acc0[i][j] = fmaf(a0.z, fz, acc0[i][j]); acc1[i][j] = fmaf(a1.z, fz, acc1[i][j]); }
#pragma unroll
                for (int j = 0; j < CV_D; ++j) { const float fw = f[(i + j) % CV_D].w; acc0[i][j] = fmaf(a0.w, fw, acc0[i][j]); acc1[i][j] = fmaf(a1.w, fw, acc1[i][j]); }
                if (i + 1 < CV_SX)   // column i leaves the window, column i + 9 enters
                    f[i % CV_D] = *reinterpret_cast<const float4*>(f_base + (i + CV_D) * CV_CH + 4 * k);
            }
        }
    }
    __syncthreads();   // everyone is done reading f1s -> reuse it as the output staging tile
    float* outs = f1s;
    const int iv0 = r_idx - ya;   // vertical shift index (v + 4) of row ya; row ya + 1 has iv0 - 1
    if (n_rows >= 1) {
#pragma unroll
        for (int i = 0; i < CV_SX; ++i)
#pragma unroll
            for (int j = 0; j < CV_D; ++j)
                outs[ya * CV_OROW + (xs * CV_SX + i) * CV_OPIX + iv0 * CV_D + j] = leaky(acc0[i][j] * p.inv_c, p.alpha);
    }
    if (n_rows == 2) {
#pragma unroll
        for (int i = 0; i < CV_SX; ++i)
#pragma unroll
            for (int j = 0; j < CV_D; ++j)
                outs[(ya + 1) * CV_OROW + (xs * CV_SX + i) * CV_OPIX + (iv0 - 1) * CV_D + j] = leaky(acc1[i][j] * p.inv_c, p.alpha);
    }
    __syncthreads();
    // ---- contiguous 324-byte runs per pixel
    const bool vec = ((p.out_cs & 3) == 0) && aligned16(p.out);
    constexpr int UNITS = 21;   // 20 float4 + 1 scalar
    for (int e = tid; e < CV_TY * CV_TX * UNITS; e += CV_THREADS) {
        const int pix = e / UNITS, u = e - pix * UNITS;
        const int py = pix / CV_TX, px = pix % CV_TX;
        const int gy = y0 + py, gx = x0 + px;
        if (gy >= p.H || gx >= p.W) continue;
        const float* s = outs + py * CV_OROW + px * CV_OPIX + 4 * u;
        float* g = p.out + (((size_t)b * p.H + gy) * p.W + gx) * p.out_cs + 4 * u;
        if (u < 20) {
            const float4 v = *reinterpret_cast<const float4*>(s);
            if (vec) *reinterpret_cast<float4*>(g) = v;
            else { g[0] = v.x; g[1] = v.y; g[2] = v.z; g[3] = v.w; }
        } else {
            g[0] = s[0];
        }
    }
}

// Generic search range (reference ctor arg search_range, model.py:88): one thread per output
// value.  Only used when search_range != 4; not tuned.
__global__ void cost_volume_generic_kernel(const CvParams p, int r) {
    const int d = 2 * r + 1, nd = d * d;
    const size_t total = (size_t)p.B * p.H * p.W * nd;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int dd = idx % nd;
        size_t pix = idx / nd;
        const int x = pix % p.W; pix /= p.W;
        const int y = pix % p.H; const int b = pix / p.H;
        const int v = dd / d - r, h = dd % d - r;
        const int yy = y + v, xx = x + h;
        float s = 0.f;
        if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) {
            const float* a = p.f0 + (((size_t)b * p.H + y) * p.W + x) * p.f0_cs;
            const float* f1b = p.f1 + (size_t)b * p.H * p.W * p.f1_cs;
            const float* flowb = p.flow ? p.flow + (size_t)b * p.H * p.W * p.flow_cs : nullptr;
            for (int c = 0; c < p.C; c += 4) {
                const float4 av = ldg4(a + c);
                float4 bv;
                if (!p.flow) bv = sample_f1<0>(p, f1b, flowb, yy, xx, c);
                else if (p.warp_type == 0) bv = sample_f1<1>(p, f1b, flowb, yy, xx, c);
                else bv = sample_f1<2>(p, f1b, flowb, yy, xx, c);
                s = fmaf(av.x, bv.x, s); s = fmaf(av.y, bv.y, s); s = fmaf(av.z, bv.z, s); s = fmaf(av.w, bv.w, s);
            }
        }
        p.out[(((size_t)b * p.H + y) * p.W + x) * p.out_cs + dd] = leaky(s * p.inv_c, p.alpha);
    }
}

__global__ void copy_channels_kernel(const float* src, int src_cs, float* dst, int dst_cs, size_t npix, int C4) {
    const size_t total = npix * C4;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t pix = idx / C4; const int k = idx % C4;
        *reinterpret_cast<float4*>(dst + pix * dst_cs + 4 * k) = ldg4(src + pix * src_cs + 4 * k);
    }
}

static int launch_cv(CvParams p, int search_range, cudaStream_t st) {
    PWC_REQUIRE(p.f0 && p.f1 && p.out, PWC_E_BADARG, "cost_volume: null pointer");
    PWC_REQUIRE(p.B > 0 && p.H > 0 && p.W > 0 && p.C > 0 && search_range >= 0, PWC_E_BADARG, "cost_volume: bad dims");
    PWC_REQUIRE(p.B <= 65535, PWC_E_BADARG, "cost_volume: batch > 65535");
    PWC_REQUIRE((p.C & 3) == 0 && (p.f0_cs & 3) == 0 && (p.f1_cs & 3) == 0 && aligned16(p.f0) && aligned16(p.f1),
                PWC_E_ALIGN, "cost_volume: C, f0_cs, f1_cs must be multiples of 4 and f0/f1 16-byte aligned");
    PWC_REQUIRE(p.f0_cs >= p.C && p.f1_cs >= p.C, PWC_E_BADARG, "cost_volume: channel stride < C");
    if (p.f0_copy)
        PWC_REQUIRE((p.f0_copy_cs & 3) == 0 && aligned16(p.f0_copy), PWC_E_ALIGN, "cost_volume: f0_copy alignment");
    p.inv_c = 1.0f / (float)p.C;
    const int nd = (2 * search_range + 1) * (2 * search_range + 1);
    PWC_REQUIRE(p.out_cs >= nd, PWC_E_BADARG, "cost_volume: out_cs < (2r+1)^2");
    if (search_range == CV_R && !p.flow && !getenv("PWC_CV_LEGACY")) {
        // PWC_CV_KERNEL=tc selects the tcgen05 band-GEMM kernel (cost_volume_tc.cu, C % 32 == 0): parity-green but
        // measured slower than the TMA-pipelined CUDA-core kernel (77-109 vs 57 us at level 2, B = 8) because it is
        // bound by shared-memory bandwidth; default = tma.  Each falls through when it does not fit.
        const char* sel = getenv("PWC_CV_KERNEL");
        if (sel && sel[0] == 't' && sel[1] == 'c') {
            const int rc = launch_cv_tc(p, st);
            if (rc != CV_TMA_UNSUPPORTED) return rc;
        }
        const int rc = launch_cv_tma(p, st);
        if (rc != CV_TMA_UNSUPPORTED) return rc;
    }
    if (search_range == CV_R) {
        dim3 grid((p.W + CV_TX - 1) / CV_TX, (p.H + CV_TY - 1) / CV_TY, p.B);
        auto kern = !p.flow ? cost_volume_r4_kernel<0> : (p.warp_type == 0 ? cost_volume_r4_kernel<1> : cost_volume_r4_kernel<2>);
        const int smem_bytes = p.flow ? CV_SMEM_BYTES_FUSED : CV_SMEM_BYTES;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) { set_error("cost_volume: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        kern<<<grid, CV_THREADS, smem_bytes, st>>>(p);
        PWC_CHECK_LAUNCH("cost_volume_r4_kernel");
    } else {
        const size_t total = (size_t)p.B * p.H * p.W * nd;
        int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
        cost_volume_generic_kernel<<<blocks, 256, 0, st>>>(p, search_range);
        PWC_CHECK_LAUNCH("cost_volume_generic_kernel");
        if (p.f0_copy) {
            const size_t npix = (size_t)p.B * p.H * p.W;
            copy_channels_kernel<<<148 * 8, 256, 0, st>>>(p.f0, p.f0_cs, p.f0_copy, p.f0_copy_cs, npix, p.C / 4);
            PWC_CHECK_LAUNCH("copy_channels_kernel");
        }
    }
    return 0;
}

}  // namespace pwc

extern "C" int pwc_cost_volume_fwd(const float* f0, int f0_cs, const float* f1, int f1_cs,
                                   float* out, int out_cs, float* f0_copy, int f0_copy_cs,
                                   int B, int H, int W, int C, int search_range, float alpha, void* stream) {
    pwc::CvParams p{};
    p.f0 = f0; p.f1 = f1; p.flow = nullptr; p.out = out; p.f0_copy = f0_copy;
    p.f0_cs = f0_cs; p.f1_cs = f1_cs; p.flow_cs = 0; p.out_cs = out_cs; p.f0_copy_cs = f0_copy_cs;
    p.B = B; p.H = H; p.W = W; p.C = C; p.flow_scale = 1.f; p.alpha = alpha; p.warp_type = 0;
    return pwc::launch_cv(p, search_range, (cudaStream_t)stream);
}

extern "C" int pwc_warp_cost_volume_fwd(const float* f0, int f0_cs, const float* f1, int f1_cs,
                                        const float* flow, int flow_cs, float flow_scale, int warp_type,
                                        float* out, int out_cs, float* f0_copy, int f0_copy_cs,
                                        int B, int H, int W, int C, int search_range, float alpha, void* stream) {
    PWC_REQUIRE(flow != nullptr, PWC_E_BADARG, "warp_cost_volume: null flow");
    PWC_REQUIRE(warp_type == 0 || warp_type == 1, PWC_E_BADARG, "warp_cost_volume: warp_type must be 0 (bilinear) or 1 (nearest)");
    PWC_REQUIRE(flow_cs >= 2, PWC_E_BADARG, "warp_cost_volume: flow_cs < 2");
    pwc::CvParams p{};
    p.f0 = f0; p.f1 = f1; p.flow = flow; p.out = out; p.f0_copy = f0_copy;
    p.f0_cs = f0_cs; p.f1_cs = f1_cs; p.flow_cs = flow_cs; p.out_cs = out_cs; p.f0_copy_cs = f0_copy_cs;
    p.B = B; p.H = H; p.W = W; p.C = C; p.flow_scale = flow_scale; p.alpha = alpha; p.warp_type = warp_type;
    return pwc::launch_cv(p, search_range, (cudaStream_t)stream);
}
