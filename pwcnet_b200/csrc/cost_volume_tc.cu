// Cost volume, search range 4, on the tcgen05 tensor cores (sm_100a) -- the level-2 roofline kernel.
//
//   out[b,y,x,(v+4)*9+(h+4)] = leaky( (1/C) * sum_c f0[b,y,x,c] * f1[b,y+v,x+h,c] ),  zero outside the image
//   (CostVolumeLayer.__call__ / get_cost, reference modules.py:164-204).
//
// Why tensor cores.  The op moves 580 B per pixel (C = 32) but needs 2592 FMAs per pixel: on the FP32 pipe the
// arithmetic alone is 16.5 us at B = 8 against 20.3 us of HBM time, and the best CUDA-core kernel
// (cost_volume_tma.cu) sits at 0.36 of the HBM roofline, bound by instruction issue.  The correlation IS a
// (banded) contraction over channels, so it is done as a dense GEMM per tile and the band is cut out afterwards:
//
//   tile      : f0 patch 16 x 8 pixels (M = 128 rows) against the f1 patch 24 x 16 pixels around it
//               (N = 384 = two halves of 192 columns: f1 rows 0..7 and 8..15), K = 32 channels per chunk.
//   D[m, n]   = sum_c f0[m, c] * f1[n, c]; pixel m = (py, px) needs the 81 columns n = (py+v')*24 + (px+h'),
//               v', h' in 0..8: 21 % of the MMA work is used, which still is ~10x fewer issue slots than FFMA.
//   precision : "3 x fp16": x = h + l, h = fp16(x), l = fp16(x - h); D = h.h + l.h + h.l accumulated in ONE fp32
//               TMEM accumulator (h.h products are exact; l is kept unscaled -- fp16 subnormals bound its error by
//               3e-8 absolute, i.e. fp32-class for the O(1e-2..1e2) features of this network).
//   pipeline  : TMA producer warp (f0 box {32ch,16,8}, f1 box {32ch,24,16} at (x0-4, y0-4); out-of-bounds zero
//               fill IS the reference's zero padding, 128B swizzle) -> 8 converter warps split each fp32 row
//               (128 B) IN PLACE into [h: 32 x fp16 | l: 32 x fp16] (a K-major, 128B-swizzled UMMA operand row of
//               K = 64) -> one thread issues 12 tcgen05.mma (M128 N192 K16) per chunk -> 4 epilogue warps read the
//               accumulator with tcgen05.ld (lane = pixel), scatter the band into a per-warp shared-memory slab
//               [pixel][81] and write 324-byte runs to HBM.  2 smem stages of 64 KB; the two accumulator halves
//               are released to the MMA warp independently, so the MMAs of tile t+1 overlap the epilogue of tile t.
//   f0_copy   : the converter threads that own f0 rows also write them to the concat slot (modules.py:262).
#include "cost_volume.cuh"
#include "tc_common.cuh"
#include <cuda_fp16.h>
#include <cstdlib>

namespace pwc {

constexpr int X_TW = 16, X_TH = 8;                 // f0 patch
constexpr int X_FW = X_TW + 8, X_FH = X_TH + 8;    // f1 patch 24 x 16
constexpr int X_M = X_TW * X_TH;                   // 128
constexpr int X_NH = X_FW * (X_FH / 2);            // 192 columns per accumulator half
constexpr int X_BK = 32;
constexpr uint32_t X_F0_BYTES = X_M * X_BK * 4;            // 16 KB
constexpr uint32_t X_F1_BYTES = 2 * X_NH * X_BK * 4;       // 48 KB
constexpr uint32_t X_STAGE_BYTES = X_F0_BYTES + X_F1_BYTES;
constexpr int X_STAGES = 2;
constexpr int X_ROWS = X_M + 2 * X_NH;             // 512 operand rows per stage
constexpr int X_CONV_WARPS = 8, X_CONV_THREADS = X_CONV_WARPS * 32;
constexpr int X_EPI_WARPS = 8;                     // two per TMEM lane quadrant: one per accumulator half
constexpr int X_THREADS = 64 + X_EPI_WARPS * 32 + X_CONV_THREADS;   // TMA warp, MMA warp, 8 epilogue warps, 8 converter warps
constexpr int X_SLAB_PITCH = 84;                   // floats per pixel row of the slab (81 used, 16-byte aligned rows)
constexpr uint32_t X_SLAB_BYTES = 4 * 32 * X_SLAB_PITCH * 4;
constexpr uint32_t X_SMEM_BYTES = X_STAGES * X_STAGE_BYTES + X_SLAB_BYTES + 1024;
constexpr uint32_t X_TMEM_COLS = 512;

struct CvTcParams {
    float* out; float* f0_copy;
    int out_cs, f0_copy_cs, B, H, W, kchunks;
    int tiles_x, tiles_y, total_tiles;
    float alpha, inv_c;
    int vec, backoff;
    unsigned long long* dbg;   // optional timeline (clock64), 64 slots per CTA: 8 events x 8 tiles; nullptr in production
};

// Waiting warps that are not on the critical issue path back off between polls: every mbarrier.try_wait is a
// shared-memory access that competes with the tensor core's operand fetches.
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P1;\n\t"
            "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(128);
    }
}

#define X_DBG(ev, tile) do { if (dbg && (tile) < 8) dbg[(ev) * 8 + (tile)] = clock64(); } while (0)

__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// Band extraction of one accumulator half: register i of 16-column chunk j holds D[m, n], n = HB*192 + 16j + i,
// i.e. f1 patch pixel (fy, fx) = (HB*8 + (16j+i)/24, (16j+i)%24) -- compile-time after unrolling; the lane's pixel
// (py, px) decides whether it is inside the 9x9 window and where it goes in the slab row.
template <int HB>
__device__ __forceinline__ void scatter_half(uint32_t taddr, float* slab_lane, int py, int px, int q) {
    const uint32_t rowmask = 0x1FFu << py, colmask = 0x1FFu << px;
#pragma unroll
    for (int j = 0; j < X_NH / 16; ++j) {
        const int fylo = HB * 8 + (16 * j) / X_FW, fyhi = HB * 8 + (16 * j + 15) / X_FW;
        if (fyhi < 2 * q || fylo > 2 * q + 9) continue;   // no lane of this warp (py in {2q, 2q+1}) needs the chunk
        uint32_t r[16];
        tmem_ld16(taddr + 16 * j, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int n = 16 * j + i;
            const int fy = HB * 8 + n / X_FW, fx = n % X_FW;
            if (((rowmask >> fy) & (colmask >> fx)) & 1u) slab_lane[fy * 9 + fx] = __uint_as_float(r[i]);
        }
    }
}

__global__ void __launch_bounds__(X_THREADS, 1)
cost_volume_tc_kernel(const __grid_constant__ CUtensorMap tm_f0, const __grid_constant__ CUtensorMap tm_f1, const CvTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    __shared__ __align__(8) uint64_t bars[3 * X_STAGES + 4];   // full[2], conv[2], empty[2], acc_full[2], acc_empty[2]
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_full = smem_u32(&bars[0]), bar_conv = smem_u32(&bars[X_STAGES]), bar_empty = smem_u32(&bars[2 * X_STAGES]);
    const uint32_t bar_accf = smem_u32(&bars[3 * X_STAGES]), bar_acce = smem_u32(&bars[3 * X_STAGES + 2]);

    if (threadIdx.x == 0) {
        for (int s = 0; s < X_STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_conv + 8 * s, X_CONV_THREADS);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int h = 0; h < 2; ++h) {
            mbar_init(bar_accf + 8 * h, 1);
            mbar_init(bar_acce + 8 * h, 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(X_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = tmem_base_slot;
    const int KC = p.kchunks;
    unsigned long long* dbg = p.dbg ? p.dbg + (size_t)blockIdx.x * 64 : nullptr;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_f0) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_f1) : "memory");
            int it = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const int tx = t % p.tiles_x, ty = (t / p.tiles_x) % p.tiles_y, b = t / (p.tiles_x * p.tiles_y);
                const int x0 = tx * X_TW, y0 = ty * X_TH;
                for (int c = 0; c < KC; ++c, ++it) {
                    const int s = it % X_STAGES;
                    const uint32_t ph = (it / X_STAGES) & 1;
                    mbar_wait(bar_empty + 8 * s, ph ^ 1);
                    const uint32_t st = base + s * X_STAGE_BYTES;
                    X_DBG(0, it);
                    mbar_expect_tx(bar_full + 8 * s, X_STAGE_BYTES);
                    tma_load_4d(st, &tm_f0, bar_full + 8 * s, c * X_BK, x0, y0, b);
                    tma_load_4d(st + X_F0_BYTES, &tm_f1, bar_full + 8 * s, c * X_BK, x0 - 4, y0 - 4, b);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            // kind::f16: D = f32 (bit 4), A = B = F16, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
            const uint32_t idesc = (1u << 4) | ((uint32_t)(X_NH >> 3) << 17) | ((uint32_t)(X_M >> 4) << 24);
            // K-major rows of 128 bytes (64 fp16: h | l), 128B swizzle, 8-row atoms 1024 bytes apart
            const uint64_t desc_hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
            int it = 0, tcount = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++tcount) {
                for (int c = 0; c < KC; ++c, ++it) {
                    const int s = it % X_STAGES;
                    const uint32_t ph = (it / X_STAGES) & 1;
                    mbar_wait(bar_conv + 8 * s, ph);
                    X_DBG(3, it);
                    tc_fence_after();
                    const uint32_t st = base + s * X_STAGE_BYTES;
                    const uint32_t a0 = ((st >> 4) & 0x3FFF) | (1u << 16);
#pragma unroll
                    for (int hb = 0; hb < 2; ++hb) {
                        if (c == 0 && tcount > 0) {   // the epilogue of the previous tile has drained this half
                            mbar_wait(bar_acce + 8 * hb, (tcount - 1) & 1);
                            tc_fence_after();
                        }
                        const uint32_t b0 = (((st + X_F0_BYTES + hb * X_NH * 128) >> 4) & 0x3FFF) | (1u << 16);
                        const uint32_t d = tmem_acc + hb * X_NH;
                        // k-steps of 32 bytes inside the 128-byte row: 0,1 = h (channels 0-15, 16-31), 2,3 = l
#pragma unroll
                        for (int k = 0; k < 2; ++k) mma_f16_ss(d, desc_hi | (a0 + 2 * k), desc_hi | (b0 + 2 * k), idesc, (c | k) != 0 ? 1u : 0u);
#pragma unroll
                        for (int k = 0; k < 2; ++k) mma_f16_ss(d, desc_hi | (a0 + 4 + 2 * k), desc_hi | (b0 + 2 * k), idesc, 1u);
#pragma unroll
                        for (int k = 0; k < 2; ++k) mma_f16_ss(d, desc_hi | (a0 + 2 * k), desc_hi | (b0 + 4 + 2 * k), idesc, 1u);
                        if (c == KC - 1) tc_commit(bar_accf + 8 * hb);
                    }
                    tc_commit(bar_empty + 8 * s);
                    X_DBG(4, it);
                }
            }
        }
    } else if (warp < 2 + X_EPI_WARPS) {
        // ===================== epilogue (warps 2..9) =====================
        // TMEM lane quadrant q = warp % 4 (hardware rule); warps 2..5 read accumulator half A, warps 6..9 half B.
        // The two warps of a quadrant share one slab (disjoint displacement sets), pair up on a named barrier and
        // then each writes half of the quadrant's 32 pixels x 81 floats to HBM.
        const int q = warp & 3, hb = (warp - 2) >> 2;
        const int py = 2 * q + (lane >> 4), px = lane & 15;
        float* slab = reinterpret_cast<float*>(base_ptr + X_STAGES * X_STAGE_BYTES) + q * 32 * X_SLAB_PITCH;
        float* slab_lane = slab + lane * X_SLAB_PITCH - (py * 9 + px);
        const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + hb * X_NH;
        const int cs = p.out_cs;
        const float inv_c = p.inv_c, alpha = p.alpha;
        int tcount = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++tcount) {
            const int tx = t % p.tiles_x, ty = (t / p.tiles_x) % p.tiles_y, b = t / (p.tiles_x * p.tiles_y);
            const int x0 = tx * X_TW, y0 = ty * X_TH;
            if (p.backoff) mbar_wait_backoff(bar_accf + 8 * hb, tcount & 1); else mbar_wait(bar_accf + 8 * hb, tcount & 1);
            if (lane == 0) X_DBG(5 + hb, tcount);
            tc_fence_after();
            if (hb == 0) scatter_half<0>(taddr, slab_lane, py, px, q);
            else scatter_half<1>(taddr, slab_lane, py, px, q);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acce + 8 * hb);
            asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");      // both halves of the band are in the slab
            // ---- slab -> HBM: the quadrant owns image rows y0+2q, y0+2q+1 (16 pixels each, 81 floats per pixel);
            // this warp writes pixel row `hb` of the two.
            const int yy = y0 + 2 * q + hb;
            if (yy < p.H) {
                float* orow = p.out + (((size_t)b * p.H + yy) * p.W + x0) * cs;
                const float* srow = slab + hb * 16 * X_SLAB_PITCH;
                const int npx = min(X_TW, p.W - x0);
                if (p.vec) {
                    // 16 pixels x (20 float4 + 1 scalar) = 336 units; lane handles u = lane + 32 m
                    int pix = lane / 21, k = lane - pix * 21;
#pragma unroll
                    for (int m = 0; m < 11; ++m) {
                        if (pix < npx) {
                            const float* src = srow + pix * X_SLAB_PITCH + 4 * k;
                            float* dst = orow + pix * cs + 4 * k;
                            if (k < 20) {
                                float4 v = *reinterpret_cast<const float4*>(src);
                                v.x *= inv_c; v.y *= inv_c; v.z *= inv_c; v.w *= inv_c;
                                v.x = fmaxf(v.x, alpha * v.x); v.y = fmaxf(v.y, alpha * v.y);
                                v.z = fmaxf(v.z, alpha * v.z); v.w = fmaxf(v.w, alpha * v.w);
                                *reinterpret_cast<float4*>(dst) = v;
                            } else {
                                const float v = *src * inv_c;
                                *dst = fmaxf(v, alpha * v);
                            }
                        }
                        k += 11; pix += 1;                  // u += 32 = 21 + 11
                        if (k >= 21) { k -= 21; pix += 1; }
                    }
                } else {
                    // 16 pixels x 81 scalars = 1296 units
                    int pix = lane / 81, k = lane - pix * 81;   // = 0, lane
#pragma unroll 3
                    for (int m = 0; m < 41; ++m) {
                        if (pix < npx) {
                            const float v = srow[pix * X_SLAB_PITCH + k] * inv_c;
                            orow[pix * cs + k] = fmaxf(v, alpha * v);
                        }
                        k += 32;
                        if (k >= 81) { k -= 81; pix += 1; }
                    }
                }
            }
            asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");      // the slab is rewritten by the next tile's scatter
            if (warp == 2 && lane == 0) X_DBG(7, tcount);
        }
    } else {
        // ===================== converters (warps 10..17): fp32 row -> [h | l] fp16 row, in place =====================
        const int ct = threadIdx.x - (64 + X_EPI_WARPS * 32);   // 0..255
        int it = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            const int tx = t % p.tiles_x, ty = (t / p.tiles_x) % p.tiles_y, b = t / (p.tiles_x * p.tiles_y);
            const int x0 = tx * X_TW, y0 = ty * X_TH;
            for (int c = 0; c < KC; ++c, ++it) {
                const int s = it % X_STAGES;
                const uint32_t ph = (it / X_STAGES) & 1;
                if (p.backoff) mbar_wait_backoff(bar_full + 8 * s, ph); else mbar_wait(bar_full + 8 * s, ph);
                if (ct == 0) X_DBG(1, it);
                uint8_t* stp = base_ptr + (size_t)s * X_STAGE_BYTES;
#pragma unroll
                for (int rr = 0; rr < X_ROWS / X_CONV_THREADS; ++rr) {
                    const int R = ct + rr * X_CONV_THREADS;    // operand row: f0 rows 0..127, then f1 rows 0..383
                    uint8_t* row = stp + (size_t)R * 128;
                    const int sw = R & 7;                      // 128B swizzle: logical 16-byte chunk j sits at chunk j ^ (R & 7)
                    float4 v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = *reinterpret_cast<const float4*>(row + ((j ^ sw) << 4));
                    if (rr == 0 && ct < X_M && p.f0_copy) {    // f0 rows: also the estimator's concat slot
                        const int yy = y0 + (ct >> 4), xx = x0 + (ct & 15);
                        if (yy < p.H && xx < p.W) {
                            float* dst = p.f0_copy + (((size_t)b * p.H + yy) * p.W + xx) * p.f0_copy_cs + c * X_BK;
#pragma unroll
                            for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(dst + 4 * j) = v[j];
                        }
                    }
                    uint4 hq[4], lq[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 a = v[2 * j], bq = v[2 * j + 1];
                        const __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
                        const __half2 h2 = __floats2half2_rn(bq.x, bq.y), h3 = __floats2half2_rn(bq.z, bq.w);
                        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1), f2 = __half22float2(h2), f3 = __half22float2(h3);
                        const __half2 l0 = __floats2half2_rn(a.x - f0.x, a.y - f0.y), l1 = __floats2half2_rn(a.z - f1.x, a.w - f1.y);
                        const __half2 l2 = __floats2half2_rn(bq.x - f2.x, bq.y - f2.y), l3 = __floats2half2_rn(bq.z - f3.x, bq.w - f3.y);
                        hq[j] = make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                                           *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
                        lq[j] = make_uint4(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1),
                                           *reinterpret_cast<const uint32_t*>(&l2), *reinterpret_cast<const uint32_t*>(&l3));
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        *reinterpret_cast<uint4*>(row + ((j ^ sw) << 4)) = hq[j];
                        *reinterpret_cast<uint4*>(row + (((j + 4) ^ sw) << 4)) = lq[j];
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                if (ct == 0) X_DBG(2, it);
                mbar_arrive(bar_conv + 8 * s);
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(X_TMEM_COLS));
    }
}

static bool make_map_tc(CUtensorMap* tm, const float* basep, int cs, int B, int H, int W, int C, int bx, int by) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)cs * 4, (cuuint64_t)W * cs * 4, (cuuint64_t)H * W * cs * 4};
    const cuuint32_t box[4] = {(cuuint32_t)X_BK, (cuuint32_t)bx, (cuuint32_t)by, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(basep), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int launch_cv_tc(const CvParams& q, cudaStream_t st) {
    if (q.flow || (q.C % X_BK) != 0 || (q.f0_cs & 3) || (q.f1_cs & 3) || !aligned16(q.f0) || !aligned16(q.f1))
        return CV_TMA_UNSUPPORTED;
    if (q.f0_copy && ((q.f0_copy_cs & 3) || !aligned16(q.f0_copy))) return CV_TMA_UNSUPPORTED;
    CUtensorMap tm0, tm1;
    if (!make_map_tc(&tm0, q.f0, q.f0_cs, q.B, q.H, q.W, q.C, X_TW, X_TH) ||
        !make_map_tc(&tm1, q.f1, q.f1_cs, q.B, q.H, q.W, q.C, X_FW, X_FH))
        return CV_TMA_UNSUPPORTED;
    CvTcParams p{};
    p.out = q.out; p.f0_copy = q.f0_copy; p.out_cs = q.out_cs; p.f0_copy_cs = q.f0_copy_cs;
    p.B = q.B; p.H = q.H; p.W = q.W; p.kchunks = q.C / X_BK;
    p.tiles_x = (q.W + X_TW - 1) / X_TW; p.tiles_y = (q.H + X_TH - 1) / X_TH;
    const long long tiles = (long long)p.tiles_x * p.tiles_y * q.B;
    if (tiles >= (1ll << 30)) return CV_TMA_UNSUPPORTED;
    p.total_tiles = (int)tiles;
    p.alpha = q.alpha; p.inv_c = q.inv_c;
    p.vec = aligned16(q.out) && (q.out_cs & 3) == 0;
    p.backoff = getenv("PWC_CV_NOBACKOFF") ? 0 : 1;
    cudaError_t e = cudaFuncSetAttribute(cost_volume_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)X_SMEM_BYTES);
    if (e != cudaSuccess) { set_error("cost_volume_tc: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    const int nsm = sm_count();
    const int grid = p.total_tiles < nsm ? p.total_tiles : nsm;
    static unsigned long long* dbg_buf = nullptr;
    if (getenv("PWC_CV_DEBUG")) {
        if (!dbg_buf) cudaMalloc(&dbg_buf, 256 * 64 * 8);
        cudaMemsetAsync(dbg_buf, 0, 256 * 64 * 8, st);
        p.dbg = dbg_buf;
    }
    cost_volume_tc_kernel<<<grid, X_THREADS, X_SMEM_BYTES, st>>>(tm0, tm1, p);
    PWC_CHECK_LAUNCH("cost_volume_tc_kernel");
    if (p.dbg) {   // debugging aid only (synchronises): timeline of the first tiles of two CTAs
        cudaStreamSynchronize(st);
        static int printed = 0;
        if (printed++ < 2) {
            unsigned long long h[64];
            const int ids[2] = {0, grid / 2};
            const char* names[8] = {"tma_issue", "full_seen", "conv_done", "mma_start", "mma_issued", "accA_seen", "accB_seen", "stored"};
            for (int i = 0; i < 2; ++i) {
                cudaMemcpy(h, p.dbg + 64 * ids[i], 64 * 8, cudaMemcpyDeviceToHost);
                fprintf(stderr, "[cv_tc dbg] cta %d (clk from first tma issue), tiles 0..7\n", ids[i]);
                for (int e = 0; e < 8; ++e) {
                    fprintf(stderr, "   %-10s", names[e]);
                    for (int t = 0; t < 8; ++t) fprintf(stderr, " %7lld", (long long)(h[e * 8 + t] - h[0]));
                    fprintf(stderr, "\n");
                }
            }
        }
    }
    return 0;
}

}  // namespace pwc
