// Cost volume, search range 4, on tcgen05 from PRE-SPLIT fp16 operands (sm_100a) -- the pipeline's level kernel.
//
//   out[b,y,x,(v+4)*9+(h+4)] = leaky( (1/C) * sum_c f0[b,y,x,c] * f1[b,y+v,x+h,c] ),  zero outside the image
//   (CostVolumeLayer.__call__ / get_cost, reference modules.py:164-204).
//
// Same band GEMM as cost_volume_tc.cu (f0 patch 16 x 8 = 128 rows against the 24 x 16 f1 patch = 2 x 192 columns,
// 3 x fp16 split x = h + l in ONE fp32 TMEM accumulator), but the operands arrive already split: the producers
// (pwc_split_f16_fwd for pyramid features, pwc_warp_split_fwd for warped features) write, per pixel and 32-channel
// slice, the 128-byte row [h: 32 x fp16 | l: 32 x fp16] -- the same bytes as the fp32 row, and exactly the K-major
// UMMA operand row.  That removes the in-kernel converter pass, which was a third of the shared-memory traffic and
// competed with the tensor core's operand fetches (profiles/r01_cost_volume_tc_timeline.log).  Pipeline:
//   warp 0   TMA producer: f0 box {64 x fp16, 16, 8}, f1 box {64 x fp16, 24, 16} at (x0-4, y0-4), 128B swizzle,
//            out-of-bounds zero fill = the reference's zero padding; 2 stages of 64 KB
//   warp 1   MMA issuer: per slice 2 halves x 6 x tcgen05.mma (M128 N192 K16): h.h + l.h + h.l
//   warps 2..17  epilogue: four warps per TMEM lane quadrant (2 accumulator halves x 2 column ranges) read the
//            accumulator (lane = pixel), scatter the 9 x 9 band into the quadrant's shared-memory slab
//            [32 pixels][81], then each warp streams 8 pixels x 324 bytes to HBM.  A half is handed back to the MMA
//            warp as soon as its 8 warps have drained it.
// Measured (profiles/r01_cost_volume_split.log): 54 us at B = 8 (0.38 of the HBM roofline), 169 us at B = 32 (0.48).
// The MMAs are no longer the limit (1/3 of them changes nothing); the band scatter is: it is instruction-issue
// bound at ~3 instructions per accumulator element read (60 x 16 x 32 lane-elements per tile, 21 % useful), and the
// accumulator cannot be double-buffered (2 x 384 columns > 512).  Variants that were built and measured slower:
// three 128-column pieces over a 4-slot TMEM ring with scale/leaky in the scatter (76 us), per-pixel bulk copies
// (76 us) or one TMA tensor store per quadrant (126 us) for the output -- stores queue behind the tile loads in the
// SM's TMA engine.
#include "cost_volume.cuh"
#include "tc_common.cuh"
#include <cuda_fp16.h>
#include <cstdlib>
#include <cstring>

namespace pwc {

constexpr int S_TW = 16, S_TH = 8, S_FW = S_TW + 8, S_FH = S_TH + 8;
constexpr int S_M = S_TW * S_TH;                   // 128
constexpr int S_NH = S_FW * (S_FH / 2);            // 192 columns per accumulator half
constexpr uint32_t S_F0_BYTES = S_M * 128;         // 16 KB
constexpr uint32_t S_F1_BYTES = 2 * S_NH * 128;    // 48 KB
constexpr uint32_t S_STAGE_BYTES = S_F0_BYTES + S_F1_BYTES;
constexpr int S_STAGES = 2;
constexpr int S_EPI_WARPS = 16;
constexpr int S_THREADS = 64 + S_EPI_WARPS * 32;   // 576
constexpr int S_SLAB_PITCH = 84;
constexpr uint32_t S_SLAB_BYTES = 4 * 32 * S_SLAB_PITCH * 4;
constexpr uint32_t S_SMEM_BYTES = S_STAGES * S_STAGE_BYTES + S_SLAB_BYTES + 1024;

struct CvSplitParams {
    float* out;
    int out_cs, B, H, W, kchunks;
    int tiles_x, tiles_y, total_tiles;
    float alpha, scale;
    int vec;
    unsigned long long* dbg;   // optional timeline (clock64), 64 slots per CTA: 8 events x 8 tiles
    int exp;                   // experiments (PWC_CV_EXP, wrong results): 1 = no global stores, 2 = no copy-out at all
    int tma_out;               // quad kernel, slot mode: the slab quadrants leave through TMA tensor stores (tm_out)
    int wide;                  // quad kernel: also write words [81, 88) of every pixel = [tail 2 | zeros 5] (whole 32-byte sectors)
    const float* tail;         // wide: dense (B,H,W,2) source of words 81, 82 (the up-sampled flow); NULL = zeros
};

#define S_DBG(ev, tile) do { if (dbg && (tile) < 8) dbg[(ev) * 8 + (tile)] = clock64(); } while (0)

__device__ __forceinline__ void s_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// Band extraction of six 16-column chunks [J0, J0+6) of accumulator half HB: register i of chunk j holds D[m, n],
// n = HB*192 + 16j + i, i.e. f1 patch pixel (fy, fx) = (HB*8 + (16j+i)/24, (16j+i)%24) -- compile-time after
// unrolling; the lane's pixel (py, px) decides whether it is inside the 9x9 window and where it goes in the slab row.
// Kept to three instructions per element: measured, this scatter is instruction-issue bound (applying scale / leaky
// here instead of in the output loop made the kernel 2.4x slower).
template <int HB, int J0>
__device__ __forceinline__ void s_scatter(uint32_t taddr, float* slab_lane, int py, int px, int q) {
    const uint32_t rowmask = 0x1FFu << py, colmask = 0x1FFu << px;
#pragma unroll
    for (int j = J0; j < J0 + 6; ++j) {
        const int fylo = HB * 8 + (16 * j) / S_FW, fyhi = HB * 8 + (16 * j + 15) / S_FW;
        if (fyhi < 2 * q || fylo > 2 * q + 9) continue;   // no lane of this warp (py in {2q, 2q+1}) needs the chunk
        uint32_t r[16];
        tmem_ld16(taddr + 16 * j, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int n = 16 * j + i;
            const int fy = HB * 8 + n / S_FW, fx = n % S_FW;
            if (((rowmask >> fy) & (colmask >> fx)) & 1u) slab_lane[fy * 9 + fx] = __uint_as_float(r[i]);
        }
    }
}

__global__ void __launch_bounds__(S_THREADS, 1)
cost_volume_split_kernel(const __grid_constant__ CUtensorMap tm_f0, const __grid_constant__ CUtensorMap tm_f1, const CvSplitParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    __shared__ __align__(8) uint64_t bars[2 * S_STAGES + 4];   // full[2], empty[2], acc_full[2], acc_empty[2]
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[S_STAGES]);
    const uint32_t bar_accf = smem_u32(&bars[2 * S_STAGES]), bar_acce = smem_u32(&bars[2 * S_STAGES + 2]);

    if (threadIdx.x == 0) {
        for (int s = 0; s < S_STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int h = 0; h < 2; ++h) {
            mbar_init(bar_accf + 8 * h, 1);
            mbar_init(bar_acce + 8 * h, 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512));
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
                const int x0 = tx * S_TW, y0 = ty * S_TH;
                for (int c = 0; c < KC; ++c, ++it) {
                    const int s = it % S_STAGES;
                    mbar_wait(bar_empty + 8 * s, ((it / S_STAGES) & 1) ^ 1);
                    S_DBG(0, it);
                    const uint32_t st = base + s * S_STAGE_BYTES;
                    mbar_expect_tx(bar_full + 8 * s, S_STAGE_BYTES);
                    tma_load_4d(st, &tm_f0, bar_full + 8 * s, c * 64, x0, y0, b);
                    tma_load_4d(st + S_F0_BYTES, &tm_f1, bar_full + 8 * s, c * 64, x0 - 4, y0 - 4, b);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(S_NH >> 3) << 17) | ((uint32_t)(S_M >> 4) << 24);
            const uint64_t desc_hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
            int it = 0, tcount = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++tcount) {
                for (int c = 0; c < KC; ++c, ++it) {
                    const int s = it % S_STAGES;
                    mbar_wait(bar_full + 8 * s, (it / S_STAGES) & 1);
                    S_DBG(1, it);
                    tc_fence_after();
                    const uint32_t st = base + s * S_STAGE_BYTES;
                    const uint32_t a0 = ((st >> 4) & 0x3FFF) | (1u << 16);
#pragma unroll
                    for (int hb = 0; hb < 2; ++hb) {
                        if (c == 0 && tcount > 0) {   // the epilogue of the previous tile has drained this half
                            mbar_wait(bar_acce + 8 * hb, (tcount - 1) & 1);
                            tc_fence_after();
                        }
                        if (hb == 1) S_DBG(2, it);
                        const uint32_t b0 = (((st + S_F0_BYTES + hb * S_NH * 128) >> 4) & 0x3FFF) | (1u << 16);
                        const uint32_t d = tmem_acc + hb * S_NH;
                        // k-steps of 32 bytes inside the 128-byte row: 0,1 = h (channels 0-15, 16-31), 2,3 = l
#pragma unroll
                        for (int k = 0; k < 2; ++k) s_mma(d, desc_hi | (a0 + 2 * k), desc_hi | (b0 + 2 * k), idesc, (c | k) != 0 ? 1u : 0u);
#pragma unroll
                        for (int k = 0; k < 2; ++k) s_mma(d, desc_hi | (a0 + 4 + 2 * k), desc_hi | (b0 + 2 * k), idesc, 1u);
#pragma unroll
                        for (int k = 0; k < 2; ++k) s_mma(d, desc_hi | (a0 + 2 * k), desc_hi | (b0 + 4 + 2 * k), idesc, 1u);
                        if (c == KC - 1) tc_commit(bar_accf + 8 * hb);
                    }
                    tc_commit(bar_empty + 8 * s);
                    S_DBG(3, it);
                }
            }
        }
    } else {
        // ===================== epilogue (warps 2..17) =====================
        // TMEM lane quadrant q = warp % 4 (hardware rule).  e = (warp - 2) / 4 in 0..3: accumulator half hb = e >> 1,
        // column range ch = e & 1 (16-column chunks [6 ch, 6 ch + 6) of the half).
        const int q = warp & 3, e = (warp - 2) >> 2, hb = e >> 1, ch = e & 1;
        const int py = 2 * q + (lane >> 4), px = lane & 15;
        float* slab = reinterpret_cast<float*>(base_ptr + S_STAGES * S_STAGE_BYTES) + q * 32 * S_SLAB_PITCH;
        float* slab_lane = slab + lane * S_SLAB_PITCH - (py * 9 + px);
        const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + hb * S_NH;
        const int cs = p.out_cs;
        const float scale = p.scale, alpha = p.alpha;
        int tcount = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++tcount) {
            const int tx = t % p.tiles_x, ty = (t / p.tiles_x) % p.tiles_y, b = t / (p.tiles_x * p.tiles_y);
            const int x0 = tx * S_TW, y0 = ty * S_TH;
            mbar_wait(bar_accf + 8 * hb, tcount & 1);
            if (lane == 0 && q == 2 && ch == 0) S_DBG(4 + hb, tcount);   // warps 2 (half A) and 10 (half B)
            tc_fence_after();
            if (hb == 0) { if (ch == 0) s_scatter<0, 0>(taddr, slab_lane, py, px, q); else s_scatter<0, 6>(taddr, slab_lane, py, px, q); }
            else         { if (ch == 0) s_scatter<1, 0>(taddr, slab_lane, py, px, q); else s_scatter<1, 6>(taddr, slab_lane, py, px, q); }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acce + 8 * hb);
            if (warp == 2 && lane == 0) S_DBG(6, tcount);
            asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");     // the whole band of the quadrant is in the slab
            // ---- slab -> HBM: the quadrant owns image rows y0+2q, y0+2q+1 (16 pixels each); this warp writes
            // eight pixels: row e >> 1, columns 8 (e & 1) .. +7.  scale and the leaky slope are applied here, on the 81
            // useful values per pixel.
            const int yy = y0 + 2 * q + (e >> 1), xs = x0 + 8 * (e & 1);
            if (yy < p.H && xs < p.W) {
                float* orow = p.out + (((size_t)b * p.H + yy) * p.W + xs) * cs;
                const float* srow = slab + (e * 8) * S_SLAB_PITCH;
                const int npx = min(8, p.W - xs);
                if (p.vec) {
                    // 8 pixels x (20 float4 + 1 scalar) = 168 units; lane handles u = lane + 32 m
                    int pix = lane / 21, k = lane - pix * 21;
#pragma unroll
                    for (int m = 0; m < 6; ++m) {
                        if (pix < npx) {
                            const float* src = srow + pix * S_SLAB_PITCH + 4 * k;
                            float* dst = orow + pix * cs + 4 * k;
                            if (k < 20) {
                                float4 v = *reinterpret_cast<const float4*>(src);
                                v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
                                v.x = fmaxf(v.x, alpha * v.x); v.y = fmaxf(v.y, alpha * v.y);
                                v.z = fmaxf(v.z, alpha * v.z); v.w = fmaxf(v.w, alpha * v.w);
                                *reinterpret_cast<float4*>(dst) = v;
                            } else {
                                const float v = *src * scale;
                                *dst = fmaxf(v, alpha * v);
                            }
                        }
                        k += 11; pix += 1;                  // u += 32 = 21 + 11
                        if (k >= 21) { k -= 21; pix += 1; }
                    }
                } else {
                    int pix = 0, k = lane;                  // 8 pixels x 81 scalars
#pragma unroll 3
                    for (int m = 0; m < 21; ++m) {
                        if (pix < npx) {
                            const float v = srow[pix * S_SLAB_PITCH + k] * scale;
                            orow[pix * cs + k] = fmaxf(v, alpha * v);
                        }
                        k += 32;
                        if (k >= 81) { k -= 81; pix += 1; }
                    }
                }
            }
            asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");     // the slab is rewritten by the next tile's scatter
            if (warp == 2 && lane == 0) S_DBG(7, tcount);
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(512));
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Quadrant-block kernel (round 2, the default).  Same band GEMM, re-tiled so that the band extraction
// needs no shared-memory scatter and the output leaves through dedicated store warps:
//   tile = 16 image rows x 8 pixels (M = 128, A row m = 8 y + x: ONE TMA box), f1 patch 24 rows x 16 columns = 384
//   accumulator columns n = 16 fy + fx in two halves of 192 (one box, two N = 192 MMAs per K step).
//   TMEM lane quadrant q = image rows 4q .. 4q+3 of the tile (a 4 x 8 pixel block, lane = 8 yy + xx): the block's
//   windows span patch rows 4q .. 4q+11 x all 16 columns = 192 CONTIGUOUS accumulator columns [64 q, 64 q + 192).
//   warp 0      TMA producer (2 stages of 64 KB)
//   warp 1      MMA issuer (h.h + l.h + h.l, 12 x M128 N192 K16 per 32 channels)
//   warps 2-13  extract: three per quadrant, ROLE r owns patch rows 4r .. 4r+3 (64 columns, ONE tcgen05.ld x64 straight
//               into registers; the accumulator goes back to the MMA warp as soon as the load has landed).  The
//               lane-dependent window offset is resolved without a scatter: columns by a 3-stage select network in
//               registers (shift by xx = 4/2/1: 31 SEL per row), rows by the ADDRESS of the store into the lane's slab row
//               (word 9 (J - yy) + i = lane-constant base + immediate; rows outside the window predicated off).
//               (rows outside the window predicated off); scale and the leaky slope are applied to those nine words.
//   warps 14-15 store: two quadrants each, a pure copy slab (32 pixels x 81 words, pitch 84, two buffers) -> HBM as
//               float4 units of the 324-byte pixel runs (20 float4 + the 81st word per pixel; lane <-> unit map and
//               offsets precomputed once, 10 LDS.128 in flight per batch).
// Variants measured slower (profiles/r02_cv_quad.log): the extract warps doing the copy-out themselves (33-36 us: their
// select network is ALU-bound, the copy-out latency-bound, and the hand-back of the accumulator waits for the slowest warp),
// 96-word heads / 128-byte-aligned slots (33 us), two extract roles with x64 + x32 loads (ptxas sinks the second load below
// the select network of the first: the MMA warp waits 900 instead of 400 clk).
// Measured on B200 (profiles/r02_tmem_ld_bench.log, r02_cv_quad_timeline.log): tcgen05.ld is latency- not bandwidth-bound
// (218 clk per isolated x16 load, >450 B/clk/SM with 16 warps x 4 loads in flight); fma.rn.f32x2 issues at the FMA-pipe
// rate of scalar FFMA; 16-byte-per-lane stores of 324-byte runs are 2.2x slower than coalesced float4 units (hence the
// slab); with the extract warps also doing the copy-out a tile took 4.6 k clk of which the copy-out 3.3 k (2.2 k without
// the global stores: the LDS / FP / STG chain of a warp that owns the accumulator hand-over), 1.7 k without it.
constexpr int Q_TW = 8, Q_TH = 16, Q_FW = Q_TW + 8, Q_FH = Q_TH + 8;
constexpr int Q_M = Q_TW * Q_TH;                   // 128
constexpr int Q_NH = Q_FW * (Q_FH / 2);            // 192 columns per accumulator half
constexpr uint32_t Q_F0_BYTES = Q_M * 128;         // 16 KB
constexpr uint32_t Q_F1_BYTES = 2 * Q_NH * 128;    // 48 KB
constexpr uint32_t Q_STAGE_BYTES = Q_F0_BYTES + Q_F1_BYTES;
constexpr int Q_STAGES = 2;
constexpr int Q_ROLES = 3;                            // extract warps per quadrant, four patch rows (one tcgen05.ld x64) each
constexpr int Q_XWARPS = 4 * Q_ROLES, Q_SWARPS = 2;   // 16 warps = 4 per SM sub-partition: 128 registers per thread
constexpr int Q_THREADS = 64 + (Q_XWARPS + Q_SWARPS) * 32;   // 512
constexpr int Q_PITCH = 84;                        // slab row pitch in words (LDS.128 of a pixel run is conflict-free)
constexpr int Q_PITCH_TMA = 88;                    // TMA copy-out (slot mode): dense rows of the 88 words [cv 81 | zeros 7] that leave as one box
constexpr uint32_t Q_SLAB_BYTES = 4 * 32 * Q_PITCH_TMA * 4;   // one buffer: 4 quadrants x 32 pixels (sized for the larger pitch)
constexpr uint32_t Q_SMEM_BYTES = Q_STAGES * Q_STAGE_BYTES + 2 * Q_SLAB_BYTES + 1024;
static_assert(Q_SMEM_BYTES <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t* r) {
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
                 : "r"(taddr));
}

// Band extraction of one extract warp: ROLE r owns patch rows J = 4r .. 4r+3 of the quadrant's twelve -- ONE tcgen05.ld
// x64, so nothing can be scheduled between the TMEM reads and the hand-back of the accumulator (with two loads per warp
// ptxas sank the second one below the select network of the first and the MMA warp waited ~900 clk instead of ~400).  Row J
// holds displacement row dv = J - yy of the lane's pixel (if 0 <= dv <= 8): its 16 words are shifted left by xx in
// registers and the nine window words are scaled, passed through the leaky slope and stored to the lane's slab row at word
// 9 dv + i = (9 J + i) - 9 yy.  (Bank of the store = 20 xx - 9 yy + const mod 32: the 32 lanes hit 32 different banks.)
template <int ROLE, bool SCALED>
__device__ __forceinline__ void q_extract(uint32_t tq, int q, int lane, uint32_t* slab_lane, uint32_t bar_acce, float scale, float alpha) {
    uint32_t v[64];
    tmem_ld64(tq + (uint32_t)((4 * q + 4 * ROLE) * Q_FW), v);
    tmem_ld_wait();
    // the accumulator words are in registers: hand this warp's share of the accumulator back to the MMA warp
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_acce);
    const bool c4 = (lane & 4) != 0, c2 = (lane & 2) != 0, c1 = (lane & 1) != 0;      // xx = lane & 7
    const int yy = lane >> 3;
    uint32_t* dst = slab_lane - 9 * yy;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int J = 4 * ROLE + j;
        uint32_t a[12], b[10], w[9];
#pragma unroll
        for (int i = 0; i < 12; ++i) a[i] = c4 ? v[16 * j + i + 4] : v[16 * j + i];
#pragma unroll
        for (int i = 0; i < 10; ++i) b[i] = c2 ? a[i + 2] : a[i];
#pragma unroll
        for (int i = 0; i < 9; ++i) w[i] = c1 ? b[i + 1] : b[i];
        if ((unsigned)(J - yy) <= 8u) {
            // scale (1/C unless folded into the f0 operand) and the leaky slope, on the nine useful words only
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                float x = __uint_as_float(w[i]);
                if (SCALED) x *= scale;
                dst[9 * J + i] = __float_as_uint(fmaxf(x, alpha * x));
            }
        }
    }
}

__global__ void __launch_bounds__(Q_THREADS, 1)
cost_volume_quad_kernel(const __grid_constant__ CUtensorMap tm_f0, const __grid_constant__ CUtensorMap tm_f1,
                        const __grid_constant__ CUtensorMap tm_out, const CvSplitParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    // full[2], empty[2], acc_full, acc_empty, slab_full[4 quadrants][2 buffers], slab_empty[4][2]
    __shared__ __align__(8) uint64_t bars[2 * Q_STAGES + 2 + 16];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[Q_STAGES]);
    const uint32_t bar_accf = smem_u32(&bars[2 * Q_STAGES]), bar_acce = smem_u32(&bars[2 * Q_STAGES + 1]);
    const uint32_t bar_sfull = smem_u32(&bars[2 * Q_STAGES + 2]), bar_sempty = smem_u32(&bars[2 * Q_STAGES + 10]);

    if (threadIdx.x == 0) {
        for (int s = 0; s < Q_STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_accf, 1);
        mbar_init(bar_acce, Q_XWARPS);
        for (int i = 0; i < 8; ++i) {
            mbar_init(bar_sfull + 8 * i, Q_ROLES);   // the extract warps of the quadrant
            mbar_init(bar_sempty + 8 * i, 1);     // its store warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    const int pitch = p.tma_out ? Q_PITCH_TMA : Q_PITCH;
    if (p.tma_out) {
        // words 81..87 of every slab row are the zeros of the 88-word head; the extract warps only ever write words 0..80
        uint32_t* sl = reinterpret_cast<uint32_t*>(base_ptr + Q_STAGES * Q_STAGE_BYTES);
        for (int i = threadIdx.x; i < 2 * 128 * 7; i += Q_THREADS) sl[(i / 7) * Q_PITCH_TMA + 81 + i % 7] = 0u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = tmem_base_slot;
    const int KC = p.kchunks;
    unsigned long long* dbg = p.dbg ? p.dbg + (size_t)blockIdx.x * 64 : nullptr;
    pdl_trigger();                                             // (common.cuh) the next kernel may start its prologue

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_f0) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_f1) : "memory");
            pdl_wait();                                        // the split operands come from the previous kernels
            int it = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const int tx = t % p.tiles_x, ty = (t / p.tiles_x) % p.tiles_y, b = t / (p.tiles_x * p.tiles_y);
                const int x0 = tx * Q_TW, y0 = ty * Q_TH;
                for (int c = 0; c < KC; ++c, ++it) {
                    const int s = it % Q_STAGES;
                    mbar_wait(bar_empty + 8 * s, ((it / Q_STAGES) & 1) ^ 1);
                    S_DBG(0, it);
                    const uint32_t st = base + s * Q_STAGE_BYTES;
                    mbar_expect_tx(bar_full + 8 * s, Q_STAGE_BYTES);
                    tma_load_4d(st, &tm_f0, bar_full + 8 * s, c * 64, x0, y0, b);
                    tma_load_4d(st + Q_F0_BYTES, &tm_f1, bar_full + 8 * s, c * 64, x0 - 4, y0 - 4, b);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(Q_NH >> 3) << 17) | ((uint32_t)(Q_M >> 4) << 24);
            const uint64_t desc_hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
            int it = 0, tcount = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++tcount) {
                for (int c = 0; c < KC; ++c, ++it) {
                    const int s = it % Q_STAGES;
                    mbar_wait(bar_full + 8 * s, (it / Q_STAGES) & 1);
                    S_DBG(1, it);
                    if (c == 0 && tcount > 0) mbar_wait(bar_acce, (tcount - 1) & 1);   // all extract warps hold the previous tile in registers
                    S_DBG(2, it);
                    tc_fence_after();
                    const uint32_t st = base + s * Q_STAGE_BYTES;
                    const uint32_t a0 = ((st >> 4) & 0x3FFF) | (1u << 16);
#pragma unroll
                    for (int hb = 0; hb < 2; ++hb) {
                        const uint32_t b0 = (((st + Q_F0_BYTES + hb * Q_NH * 128) >> 4) & 0x3FFF) | (1u << 16);
                        const uint32_t d = tmem_acc + hb * Q_NH;
                        // k-steps of 32 bytes inside the 128-byte row: 0,1 = h (channels 0-15, 16-31), 2,3 = l
#pragma unroll
                        for (int k = 0; k < 2; ++k) s_mma(d, desc_hi | (a0 + 2 * k), desc_hi | (b0 + 2 * k), idesc, (c | k) != 0 ? 1u : 0u);
#pragma unroll
                        for (int k = 0; k < 2; ++k) s_mma(d, desc_hi | (a0 + 4 + 2 * k), desc_hi | (b0 + 2 * k), idesc, 1u);
#pragma unroll
                        for (int k = 0; k < 2; ++k) s_mma(d, desc_hi | (a0 + 2 * k), desc_hi | (b0 + 4 + 2 * k), idesc, 1u);
                    }
                    tc_commit(bar_empty + 8 * s);
                    if (c == KC - 1) tc_commit(bar_accf);
                }
            }
        }
    } else if (warp < 2 + Q_XWARPS) {
        // ===================== extract: warps 2..13, TMEM lane quadrant q = warp % 4, role = (warp - 2) / 4 ==========
        const int q = warp & 3, role = (warp - 2) >> 2;
        uint32_t* slabs = reinterpret_cast<uint32_t*>(base_ptr + Q_STAGES * Q_STAGE_BYTES);
        const uint32_t tq = tmem_acc + ((uint32_t)(q * 32) << 16);
        const float scale = p.scale, alpha = p.alpha;
        const bool scaled = scale != 1.0f;
        int tcount = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++tcount) {
            const int buf = tcount & 1;
            uint32_t* slab_lane = slabs + buf * (Q_SLAB_BYTES / 4) + (q * 32 + lane) * pitch;
            if (tcount >= 2) mbar_wait(bar_sempty + 8 * (2 * q + buf), ((tcount >> 1) - 1) & 1);   // the store warp has drained this buffer
            mbar_wait(bar_accf, tcount & 1);
            if (warp == 2 && lane == 0) S_DBG(3, tcount);
            tc_fence_after();
            if (scaled) {
                if (role == 0)      q_extract<0, true>(tq, q, lane, slab_lane, bar_acce, scale, alpha);
                else if (role == 1) q_extract<1, true>(tq, q, lane, slab_lane, bar_acce, scale, alpha);
                else                q_extract<2, true>(tq, q, lane, slab_lane, bar_acce, scale, alpha);
            } else {
                if (role == 0)      q_extract<0, false>(tq, q, lane, slab_lane, bar_acce, scale, alpha);
                else if (role == 1) q_extract<1, false>(tq, q, lane, slab_lane, bar_acce, scale, alpha);
                else                q_extract<2, false>(tq, q, lane, slab_lane, bar_acce, scale, alpha);
            }
            if (p.tma_out) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the slab is read by the TMA engine
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_sfull + 8 * (2 * q + buf));
            if (warp == 2 && lane == 0) S_DBG(4, tcount);
            if (warp == 6 && lane == 0) S_DBG(5, tcount);
        }
    } else {
        // ===================== store: the last two warps: quadrants {0, 1} and {2, 3} =====================
        // A pure copy (scale and leaky were applied by the extract warps): per quadrant 32 pixels x 20 float4 units, unit
        // u = lane + 32 m -> pixel u / 20, unit u % 20 (20 iterations, no divergence), then the 81st word of pixel = lane.
        pdl_wait();                                            // the tail operand and the output rows themselves
        if (p.tma_out) {
            // Slot mode: a quadrant's slab (32 pixels x 88 words, dense) is one {88 words, 8 pixels, 4 rows} box of the concat
            // buffer.  One thread hands it to the TMA engine, which reaches the strided-slot write pattern's ceiling
            // (~16 B/clk/SM, tools/store_bw_bench.cu) that two warps of LDS + STG could not (~10.5 B/clk/SM measured in this
            // kernel): the copy-out drops below the extraction time per tile.  Edge tiles are clipped by the tensor map.
            const uint32_t slab0 = base + Q_STAGES * Q_STAGE_BYTES;
            const int q_first = 2 * (warp - 2 - Q_XWARPS);
            int tcount = 0;
            if (lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_out) : "memory");
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++tcount) {
                const int tx = t % p.tiles_x, ty = (t / p.tiles_x) % p.tiles_y, b = t / (p.tiles_x * p.tiles_y);
                const int buf = tcount & 1;
                for (int qi = 0; qi < 2; ++qi) {
                    const int q = q_first + qi;
                    mbar_wait(bar_sfull + 8 * (2 * q + buf), (tcount >> 1) & 1);
                    if (warp == 2 + Q_XWARPS && lane == 0 && qi == 0) S_DBG(6, tcount);
                    if (lane == 0 && !(p.exp & 3)) {
                        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                                     ::"l"(&tm_out), "r"(slab0 + buf * Q_SLAB_BYTES + q * 32 * Q_PITCH_TMA * 4), "r"(0), "r"(tx * Q_TW),
                                       "r"(ty * Q_TH + 4 * q), "r"(b) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
                if (lane == 0) {
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // both quadrants have been read
                    mbar_arrive(bar_sempty + 8 * (2 * q_first + buf));
                    mbar_arrive(bar_sempty + 8 * (2 * (q_first + 1) + buf));
                }
                __syncwarp();
                if (warp == 2 + Q_XWARPS && lane == 0) S_DBG(7, tcount);
            }
            if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        } else {
        const float* slabs = reinterpret_cast<const float*>(base_ptr + Q_STAGES * Q_STAGE_BYTES);
        const int cs = p.out_cs;
        const int gpitch = p.W * cs;                          // < 2^31 / H (checked by the launcher)
        int soff[20], goff[20];                               // words into the quadrant's slab / bytes from the quadrant's first pixel
#pragma unroll
        for (int m = 0; m < 20; ++m) {
            const int u = lane + 32 * m, pix = u / 20, k = u - pix * 20;
            soff[m] = pix * Q_PITCH + 4 * k;
            goff[m] = ((pix >> 3) * gpitch + (pix & 7) * cs + 4 * k) * 4;
        }
        const int goff_t = ((lane >> 3) * gpitch + (lane & 7) * cs + 80) * 4;
        // tail words 81, 82 of pixel = lane (the up-sampled flow) of work item (tile t, quadrant q): loaded ONE ITEM AHEAD --
        // a load issued just before its own copy loop costs the store warp its full latency (41 instead of 32 us per launch)
        auto load_tail = [&](int t, int q) {
            float2 r = make_float2(0.f, 0.f);
            if (p.wide && p.tail && t < p.total_tiles) {
                const int tx = t % p.tiles_x, ty = (t / p.tiles_x) % p.tiles_y, b = t / (p.tiles_x * p.tiles_y);
                const int yy = ty * Q_TH + 4 * q + (lane >> 3), xx = tx * Q_TW + (lane & 7);
                if (yy < p.H && xx < p.W) r = __ldg(reinterpret_cast<const float2*>(p.tail) + (((size_t)b * p.H + yy) * p.W + xx));
            }
            return r;
        };
        const int q_first = 2 * (warp - 2 - Q_XWARPS);
        float2 tl_next = load_tail(blockIdx.x, q_first);
        int tcount = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++tcount) {
            const int tx = t % p.tiles_x, ty = (t / p.tiles_x) % p.tiles_y, b = t / (p.tiles_x * p.tiles_y);
            const int buf = tcount & 1;
#pragma unroll 1
            for (int qi = 0; qi < 2; ++qi) {
                const int q = q_first + qi;
                const int x0 = tx * Q_TW, y0 = ty * Q_TH + 4 * q;
                const float* slab = slabs + buf * (Q_SLAB_BYTES / 4) + q * 32 * Q_PITCH;
                char* gbase = reinterpret_cast<char*>(p.out + (((size_t)b * p.H + y0) * p.W + x0) * cs);
                const bool interior = y0 + 4 <= p.H && x0 + Q_TW <= p.W;
                const bool lane_ok = y0 + (lane >> 3) < p.H && x0 + (lane & 7) < p.W;      // pixel = lane (81st word, wide tail)
                const float2 tl = tl_next;
                tl_next = qi == 0 ? load_tail(t, q + 1) : load_tail(t + gridDim.x, q_first);
                mbar_wait(bar_sfull + 8 * (2 * q + buf), (tcount >> 1) & 1);
                if (warp == 2 + Q_XWARPS && lane == 0 && qi == 0) S_DBG(6, tcount);
                if (!(p.exp & 2)) {
                    if (p.vec && interior && !(p.exp & 1)) {
#pragma unroll
                        for (int m0 = 0; m0 < 20; m0 += 10) {
                            float4 w[10];
#pragma unroll
                            for (int i = 0; i < 10; ++i) w[i] = *reinterpret_cast<const float4*>(slab + soff[m0 + i]);
#pragma unroll
                            for (int i = 0; i < 10; ++i) *reinterpret_cast<float4*>(gbase + goff[m0 + i]) = w[i];
                        }
                        if (p.wide) {        // complete the pixel's sectors: words [80, 88) = [cv 80 | tail 2 | zeros 5]
                            *reinterpret_cast<float4*>(gbase + goff_t) = make_float4(slab[lane * Q_PITCH + 80], tl.x, tl.y, 0.f);
                            *reinterpret_cast<float4*>(gbase + goff_t + 16) = make_float4(0.f, 0.f, 0.f, 0.f);
                        } else {
                            *reinterpret_cast<float*>(gbase + goff_t) = slab[lane * Q_PITCH + 80];
                        }
                    } else if (p.vec && !(p.exp & 1)) {
#pragma unroll
                        for (int m = 0; m < 20; ++m) {              // edge tile: per-unit validity
                            const int pix = (lane + 32 * m) / 20;
                            if (y0 + (pix >> 3) < p.H && x0 + (pix & 7) < p.W)
                                *reinterpret_cast<float4*>(gbase + goff[m]) = *reinterpret_cast<const float4*>(slab + soff[m]);
                        }
                        if (lane_ok) {
                            if (p.wide) {
                                *reinterpret_cast<float4*>(gbase + goff_t) = make_float4(slab[lane * Q_PITCH + 80], tl.x, tl.y, 0.f);
                                *reinterpret_cast<float4*>(gbase + goff_t + 16) = make_float4(0.f, 0.f, 0.f, 0.f);
                            } else {
                                *reinterpret_cast<float*>(gbase + goff_t) = slab[lane * Q_PITCH + 80];
                            }
                        }
                    } else if (!(p.exp & 1)) {
#pragma unroll 1
                        for (int pix = 0; pix < 32; ++pix) {            // scalar path: destination not 16-byte aligned
                            if (y0 + (pix >> 3) >= p.H || x0 + (pix & 7) >= p.W) continue;
                            float* gp = reinterpret_cast<float*>(gbase) + (size_t)(pix >> 3) * gpitch + (pix & 7) * cs;
                            for (int k = lane; k < 81; k += 32) gp[k] = slab[pix * Q_PITCH + k];
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_sempty + 8 * (2 * q + buf));
            }
            if (warp == 2 + Q_XWARPS && lane == 0) S_DBG(7, tcount);
        }
        }   // !tma_out
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(512));
    }
}

// fp32 NHWC -> split rows: per pixel and 32-channel slice [h: 32 x fp16 | l: 32 x fp16], h = fp16(x), l = fp16(x - h);
// optionally also copies x to a second fp32 destination (the estimator's concat slot, modules.py:262).
__global__ void split_f16_kernel(const float* __restrict__ x, int x_cs, __half* __restrict__ out, float* __restrict__ copy,
                                 int copy_cs, size_t n_pix, int C, float scale) {
    pdl_wait();
    pdl_trigger();
    const int C8 = C >> 3;
    const size_t total = n_pix * C8;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int g = idx % C8; const size_t pix = idx / C8;
        float4 a = ldg4(x + pix * x_cs + 8 * g), bq = ldg4(x + pix * x_cs + 8 * g + 4);
        if (copy) {
            *reinterpret_cast<float4*>(copy + pix * copy_cs + 8 * g) = a;
            *reinterpret_cast<float4*>(copy + pix * copy_cs + 8 * g + 4) = bq;
        }
        a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;        // (the copy stays unscaled)
        bq.x *= scale; bq.y *= scale; bq.z *= scale; bq.w *= scale;
        const __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
        const __half2 h2 = __floats2half2_rn(bq.x, bq.y), h3 = __floats2half2_rn(bq.z, bq.w);
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1), f2 = __half22float2(h2), f3 = __half22float2(h3);
        const __half2 l0 = __floats2half2_rn(a.x - f0.x, a.y - f0.y), l1 = __floats2half2_rn(a.z - f1.x, a.w - f1.y);
        const __half2 l2 = __floats2half2_rn(bq.x - f2.x, bq.y - f2.y), l3 = __floats2half2_rn(bq.z - f3.x, bq.w - f3.y);
        // 8 channels = 16 bytes of h and 16 bytes of l inside slice g / 4
        __half* row = out + pix * (size_t)(2 * C) + (g >> 2) * 64 + (g & 3) * 8;
        *reinterpret_cast<uint4*>(row) = make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                                                    *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
        *reinterpret_cast<uint4*>(row + 32) = make_uint4(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1),
                                                         *reinterpret_cast<const uint32_t*>(&l2), *reinterpret_cast<const uint32_t*>(&l3));
    }
}

// WarpingLayer (modules.py:99-154) writing split rows: one thread per (pixel, 8 channels).
template <bool NEAREST>
__global__ void warp_split_kernel(const float* __restrict__ x, int x_cs, const float* __restrict__ flow, int flow_cs,
                                  float flow_scale, __half* __restrict__ out, int B, int H, int W, int C) {
    pdl_wait();
    pdl_trigger();
    const int C8 = C >> 3;
    const size_t total = (size_t)B * H * W * C8;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int g = idx % C8; const size_t pix = idx / C8;
        const int px = pix % W; const size_t rowi = pix / W;
        const int py = rowi % H; const size_t b = rowi / H;
        const float* fl = flow + pix * flow_cs;
        const float fx = __ldg(fl) * flow_scale, fy = __ldg(fl + 1) * flow_scale;
        const float* xb = x + b * H * W * x_cs + 8 * g;
        float r[8];
        if (NEAREST) {
            const int ix = min(max(px + (int)fx, 0), W - 1), iy = min(max(py + (int)fy, 0), H - 1);
            const float4 a = ldg4(xb + ((size_t)iy * W + ix) * x_cs), c = ldg4(xb + ((size_t)iy * W + ix) * x_cs + 4);
            r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = c.x; r[5] = c.y; r[6] = c.z; r[7] = c.w;
        } else {
            const float fx0 = floorf(fx), fy0 = floorf(fy), fx1 = fx0 + 1.f, fy1 = fy0 + 1.f;
            const float wl = (float)(W - 1), hl = (float)(H - 1);
            const int gy0 = (int)fminf(fmaxf((float)py + fy0, 0.f), hl), gy1 = (int)fminf(fmaxf((float)py + fy1, 0.f), hl);
            const int gx0 = (int)fminf(fmaxf((float)px + fx0, 0.f), wl), gx1 = (int)fminf(fmaxf((float)px + fx1, 0.f), wl);
            const float c00 = (fy1 - fy) * (fx1 - fx), c01 = (fy1 - fy) * (fx - fx0);
            const float c10 = (fy - fy0) * (fx1 - fx), c11 = (fy - fy0) * (fx - fx0);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const float4 a = ldg4(xb + ((size_t)gy0 * W + gx0) * x_cs + 4 * hh), bq = ldg4(xb + ((size_t)gy0 * W + gx1) * x_cs + 4 * hh);
                const float4 d = ldg4(xb + ((size_t)gy1 * W + gx0) * x_cs + 4 * hh), ee = ldg4(xb + ((size_t)gy1 * W + gx1) * x_cs + 4 * hh);
                // same expression order as warp_kernel (warp_resize_loss.cu): results are bit-identical
                r[4 * hh + 0] = c00 * a.x + c01 * bq.x + c10 * d.x + c11 * ee.x;
                r[4 * hh + 1] = c00 * a.y + c01 * bq.y + c10 * d.y + c11 * ee.y;
                r[4 * hh + 2] = c00 * a.z + c01 * bq.z + c10 * d.z + c11 * ee.z;
                r[4 * hh + 3] = c00 * a.w + c01 * bq.w + c10 * d.w + c11 * ee.w;
            }
        }
        __half2 h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            h[j] = __floats2half2_rn(r[2 * j], r[2 * j + 1]);
            const float2 f = __half22float2(h[j]);
            l[j] = __floats2half2_rn(r[2 * j] - f.x, r[2 * j + 1] - f.y);
        }
        __half* row = out + pix * (size_t)(2 * C) + (g >> 2) * 64 + (g & 3) * 8;
        *reinterpret_cast<uint4*>(row) = make_uint4(*reinterpret_cast<const uint32_t*>(&h[0]), *reinterpret_cast<const uint32_t*>(&h[1]),
                                                    *reinterpret_cast<const uint32_t*>(&h[2]), *reinterpret_cast<const uint32_t*>(&h[3]));
        *reinterpret_cast<uint4*>(row + 32) = make_uint4(*reinterpret_cast<const uint32_t*>(&l[0]), *reinterpret_cast<const uint32_t*>(&l[1]),
                                                         *reinterpret_cast<const uint32_t*>(&l[2]), *reinterpret_cast<const uint32_t*>(&l[3]));
    }
}

static bool make_map_split(CUtensorMap* tm, const void* basep, int B, int H, int W, int C, int bx, int by) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)(2 * C), (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    const cuuint32_t box[4] = {64, (cuuint32_t)bx, (cuuint32_t)by, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(basep), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace pwc

using namespace pwc;

extern "C" int pwc_split_f16_fwd(const float* x, int x_cs, void* out, float* copy, int copy_cs,
                                 long long n_pix, int C, float scale, void* stream) {
    PWC_REQUIRE(x && out && n_pix > 0 && C > 0, PWC_E_BADARG, "split_f16: bad arguments");
    PWC_REQUIRE((C % 32) == 0 && (x_cs & 3) == 0 && aligned16(x) && aligned16(out) &&
                (!copy || ((copy_cs & 3) == 0 && aligned16(copy))), PWC_E_ALIGN,
                "split_f16: C must be a multiple of 32, strides multiples of 4, pointers 16-byte aligned");
    const size_t total = (size_t)n_pix * (C / 8);
    const int blocks = (int)((total + 255) / 256 < (size_t)148 * 16 ? (total + 255) / 256 : (size_t)148 * 16);
    launch_pdl(split_f16_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, x, x_cs, (__half*)out, copy, copy_cs, (size_t)n_pix, C, scale);
    PWC_CHECK_LAUNCH("split_f16_kernel");
    return 0;
}

extern "C" int pwc_warp_split_fwd(const float* x, int x_cs, const float* flow, int flow_cs, float flow_scale,
                                  int warp_type, void* out, int B, int H, int W, int C, void* stream) {
    PWC_REQUIRE(x && flow && out, PWC_E_BADARG, "warp_split: null pointer");
    PWC_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, PWC_E_BADARG, "warp_split: bad dims");
    PWC_REQUIRE(warp_type == 0 || warp_type == 1, PWC_E_BADARG, "warp_split: warp_type must be 0 (bilinear) or 1 (nearest)");
    PWC_REQUIRE((C % 32) == 0 && (x_cs & 3) == 0 && aligned16(x) && aligned16(out), PWC_E_ALIGN,
                "warp_split: C must be a multiple of 32, x_cs a multiple of 4, x/out 16-byte aligned");
    const size_t total = (size_t)B * H * W * (C / 8);
    const int blocks = (int)((total + 255) / 256 < (size_t)148 * 16 ? (total + 255) / 256 : (size_t)148 * 16);
    if (warp_type == 1) launch_pdl(warp_split_kernel<true>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, x, x_cs, flow, flow_cs, flow_scale, (__half*)out, B, H, W, C);
    else launch_pdl(warp_split_kernel<false>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, x, x_cs, flow, flow_cs, flow_scale, (__half*)out, B, H, W, C);
    PWC_CHECK_LAUNCH("warp_split_kernel");
    return 0;
}

static int cv_split_launch(const void* f0s, const void* f1s, float* out, int out_cs, int wide, const float* tail,
                           int B, int H, int W, int C, float scale, float alpha, void* stream);

extern "C" int pwc_cost_volume_split_fwd(const void* f0s, const void* f1s, float* out, int out_cs,
                                         int B, int H, int W, int C, float scale, float alpha, void* stream) {
    return cv_split_launch(f0s, f1s, out, out_cs, 0, nullptr, B, H, W, C, scale, alpha, stream);
}

extern "C" int pwc_cost_volume_split_slot_fwd(const void* f0s, const void* f1s, float* out, int out_cs, int head, const float* tail,
                                              int B, int H, int W, int C, float scale, float alpha, void* stream) {
    PWC_REQUIRE(head == 88, PWC_E_BADARG, "cost_volume_split_slot: head must be 88");
    PWC_REQUIRE(out_cs >= 88 && (out_cs & 7) == 0 && (reinterpret_cast<uintptr_t>(out) & 31u) == 0, PWC_E_ALIGN,
                "cost_volume_split_slot: the slot must be 32-byte aligned with a pixel pitch that is a multiple of 8 floats, >= 88");
    PWC_REQUIRE(!tail || (reinterpret_cast<uintptr_t>(tail) & 7u) == 0, PWC_E_ALIGN, "cost_volume_split_slot: tail must be 8-byte aligned");
    return cv_split_launch(f0s, f1s, out, out_cs, 1, tail, B, H, W, C, scale, alpha, stream);
}

static int cv_split_launch(const void* f0s, const void* f1s, float* out, int out_cs, int wide, const float* tail,
                           int B, int H, int W, int C, float scale, float alpha, void* stream) {
    PWC_REQUIRE(f0s && f1s && out, PWC_E_BADARG, "cost_volume_split: null pointer");
    PWC_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && out_cs >= 81 && scale > 0.f, PWC_E_BADARG, "cost_volume_split: bad dims");
    PWC_REQUIRE((C % 32) == 0 && aligned16(f0s) && aligned16(f1s), PWC_E_ALIGN,
                "cost_volume_split: C must be a multiple of 32 and the operands 16-byte aligned");
    CUtensorMap tm0, tm1;
    {   // quadrant-block tiling (16 x 8 pixel tiles, register-resident band extraction): default; PWC_CV_SPLIT=scatter
        // selects the round-1 kernel below
        const char* ev = getenv("PWC_CV_SPLIT");   // read per call: the tests switch variants in one process
        if (wide || !(ev && !strcmp(ev, "scatter"))) {
            PWC_REQUIRE(make_map_split(&tm0, f0s, B, H, W, C, Q_TW, Q_TH) && make_map_split(&tm1, f1s, B, H, W, C, Q_FW, Q_FH),
                        PWC_E_BADARG, "cost_volume_split: cuTensorMapEncodeTiled failed");
            CvSplitParams p{};
            p.out = out; p.out_cs = out_cs; p.B = B; p.H = H; p.W = W; p.kchunks = C / 32;
            p.tiles_x = (W + Q_TW - 1) / Q_TW; p.tiles_y = (H + Q_TH - 1) / Q_TH;
            const long long tiles = (long long)p.tiles_x * p.tiles_y * B;
            PWC_REQUIRE(tiles < (1ll << 30), PWC_E_BADARG, "cost_volume_split: too many tiles");
            p.total_tiles = (int)tiles;
            p.alpha = alpha; p.scale = scale;
            p.vec = aligned16(out) && (out_cs & 3) == 0;
            p.wide = wide; p.tail = tail;
            // slot mode without a tail operand, 16-byte aligned rows: TMA copy-out
            CUtensorMap tm_out;
            memset(&tm_out, 0, sizeof(tm_out));
            if (wide && !tail && p.vec && !getenv("PWC_CV_NO_TMA_OUT")) {
                cuuint64_t od[4] = {88, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
                cuuint64_t os_[3] = {(cuuint64_t)out_cs * 4, (cuuint64_t)W * out_cs * 4, (cuuint64_t)H * W * out_cs * 4};
                cuuint32_t ob[4] = {88, Q_TW, 4, 1};
                cuuint32_t oe[4] = {1, 1, 1, 1};
                EncodeTiledFn enc = pwc::get_encode();
                PWC_REQUIRE(enc != nullptr, PWC_E_NOTBUILT, "cost_volume_split: cuTensorMapEncodeTiled not available from the driver");
                CUresult r = enc(&tm_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)out, od, os_, ob, oe, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                PWC_REQUIRE(r == CUDA_SUCCESS, PWC_E_BADARG, "cost_volume_split: cuTensorMapEncodeTiled(out) failed with %d", (int)r);
                p.tma_out = 1;
            }
            PWC_REQUIRE((long long)H * W * out_cs < (1ll << 31), PWC_E_BADARG, "cost_volume_split: image too large for 32-bit offsets");
            cudaError_t e = cudaFuncSetAttribute(cost_volume_quad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Q_SMEM_BYTES);
            if (e != cudaSuccess) { set_error("cost_volume_split(quad): smem attr: %s", cudaGetErrorString(e)); return (int)e; }
            const int nsm = sm_count();
            int grid = p.total_tiles < nsm ? p.total_tiles : nsm;
            if (const char* gs = getenv("PWC_CV_GRID")) {   // experiment: 0 = balanced (fewest CTAs for the same round count), n = n CTAs
                const int g = atoi(gs);
                if (g > 0) grid = g < grid ? g : grid;
                else { const int rounds = (p.total_tiles + grid - 1) / grid; grid = (p.total_tiles + rounds - 1) / rounds; }
            }
            static unsigned long long* qdbg = nullptr;
            if (const char* ex = getenv("PWC_CV_EXP")) p.exp = atoi(ex);
            if (getenv("PWC_CV_DEBUG")) {
                if (!qdbg) cudaMalloc(&qdbg, 256 * 64 * 8);
                cudaMemsetAsync(qdbg, 0, 256 * 64 * 8, (cudaStream_t)stream);
                p.dbg = qdbg;
            }
            launch_pdl(cost_volume_quad_kernel, dim3(grid), dim3(Q_THREADS), Q_SMEM_BYTES, (cudaStream_t)stream, tm0, tm1, tm_out, p);
            PWC_CHECK_LAUNCH("cost_volume_quad_kernel");
            if (p.dbg) {   // debugging aid only (synchronises): timeline of the first tiles of one CTA
                cudaStreamSynchronize((cudaStream_t)stream);
                static int printed = 0;
                if (printed++ < 2) {
                    unsigned long long h[64];
                    const char* names[8] = {"tma_issue", "full_seen", "acce_seen", "accf_seen(w2)", "extract_done(w2)", "extract_done(w6)", "slab_full(w10)", "stored(w10)"};
                    cudaMemcpy(h, p.dbg + 64 * (grid / 2), 64 * 8, cudaMemcpyDeviceToHost);
                    fprintf(stderr, "[cv_quad dbg] cta %d (clk from first tma issue), tiles 0..7\n", grid / 2);
                    for (int e = 0; e < 8; ++e) {
                        fprintf(stderr, "   %-17s", names[e]);
                        for (int t = 0; t < 8; ++t) fprintf(stderr, " %7lld", (long long)(h[e * 8 + t] - h[0]));
                        fprintf(stderr, "\n");
                    }
                }
            }
            return 0;
        }
    }
    PWC_REQUIRE(make_map_split(&tm0, f0s, B, H, W, C, S_TW, S_TH) && make_map_split(&tm1, f1s, B, H, W, C, S_FW, S_FH),
                PWC_E_BADARG, "cost_volume_split: cuTensorMapEncodeTiled failed");
    CvSplitParams p{};
    p.out = out; p.out_cs = out_cs; p.B = B; p.H = H; p.W = W; p.kchunks = C / 32;
    p.tiles_x = (W + S_TW - 1) / S_TW; p.tiles_y = (H + S_TH - 1) / S_TH;
    const long long tiles = (long long)p.tiles_x * p.tiles_y * B;
    PWC_REQUIRE(tiles < (1ll << 30), PWC_E_BADARG, "cost_volume_split: too many tiles");
    p.total_tiles = (int)tiles;
    p.alpha = alpha; p.scale = scale;
    p.vec = aligned16(out) && (out_cs & 3) == 0;
    cudaError_t e = cudaFuncSetAttribute(cost_volume_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S_SMEM_BYTES);
    if (e != cudaSuccess) { set_error("cost_volume_split: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    const int nsm = sm_count();
    const int grid = p.total_tiles < nsm ? p.total_tiles : nsm;
    static unsigned long long* dbg_buf = nullptr;
    if (getenv("PWC_CV_DEBUG")) {
        if (!dbg_buf) cudaMalloc(&dbg_buf, 256 * 64 * 8);
        cudaMemsetAsync(dbg_buf, 0, 256 * 64 * 8, (cudaStream_t)stream);
        p.dbg = dbg_buf;
    }
    cost_volume_split_kernel<<<grid, S_THREADS, S_SMEM_BYTES, (cudaStream_t)stream>>>(tm0, tm1, p);
    PWC_CHECK_LAUNCH("cost_volume_split_kernel");
    if (p.dbg) {   // debugging aid only (synchronises): timeline of the first tiles of one CTA
        cudaStreamSynchronize((cudaStream_t)stream);
        static int printed = 0;
        if (printed++ < 2) {
            unsigned long long h[64];
            const char* names[8] = {"tma_issue", "full_seen", "halfB_go", "mma_issued", "accA_seen", "accB_seen", "scatterA_done", "stored"};
            cudaMemcpy(h, p.dbg + 64 * (grid / 2), 64 * 8, cudaMemcpyDeviceToHost);
            fprintf(stderr, "[cv_split dbg] cta %d (clk from first tma issue), tiles 0..7\n", grid / 2);
            for (int e = 0; e < 8; ++e) {
                fprintf(stderr, "   %-14s", names[e]);
                for (int t = 0; t < 8; ++t) fprintf(stderr, " %7lld", (long long)(h[e * 8 + t] - h[0]));
                fprintf(stderr, "\n");
            }
        }
    }
    return 0;
}
