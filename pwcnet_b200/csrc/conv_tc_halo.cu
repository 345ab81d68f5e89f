// 3x3 stride-1 dilation-1 convolution on tcgen05 with a HALO-RESIDENT activation tile (3 x fp16 split, fp32-class).
//
// Same function and numerics as conv_tc_f16.cu (tf.layers.Conv2D(filters,(3,3),1,'same') + bias + leaky,
// modules.py:62-67, 266-274, 306-323; also Conv2DBackpropInput through pwc_conv3x3_tc_f16_dgrad), other data flow.
// conv_tc_f16.cu streams nine shifted copies of every activation tile through TMA and the fp32->fp16 converter;
// measured, it moves ~86 B/clk/SM through shared memory and is bound by exactly that (tensor pipe ~45 % busy).  Here
//   * a tile is 128 consecutive pixels of ONE image row; per 32-channel slice the producer loads the 3 x (128+2d)
//     pixel halo box once (rows y-d, y, y+d through the TMA traversal stride; 128B swizzle, out-of-bounds zero fill =
//     SAME padding), and the converter warps split each pixel row (128 B of fp32) IN PLACE into
//     [h: 32 x fp16 | l: 32 x fp16 scaled by 2^11] -- 390 rows (d = 1) instead of 9 x 128;
//   * the A operand of tap (ky, kx) is the same shared-memory tile read through a descriptor whose start address
//     is shifted by (ky*(128+2d) + kx*d) pixel rows (the swizzle is a function of the absolute shared-memory
//     address: measured, no descriptor base offset is needed);
//   * with Cout <= 128 the [W_h | W_l] weight tiles of a tap form ONE N = 2*Cout operand: A_h x [W_h|W_l] yields the
//     main and the first correction accumulator in one instruction (A_h is read once instead of twice), A_l x W_h
//     accumulates into the correction columns: 2 MMAs and 20 KB of operand reads per K = 16 step instead of 3 / 24 KB;
//   * the CTA is persistent with two accumulator sets in TMEM, so the epilogue of tile i overlaps the MMAs of
//     tile i + 1; weights stream through their own 4-deep ring from the packed images of pwc_conv3x3_pack_weights_f16.
// Shared-memory traffic per 32-channel slice of a tile drops from ~1008 KB to ~650 KB.
#include "tc_common.cuh"
#include <cuda_fp16.h>
#include <cstdlib>

namespace pwc {

constexpr int HL_M = 128;                          // output pixels per tile (one row segment)
constexpr int HL_BH = 3;                           // halo box: 3 rows (y-d, y, y+d) x (128 + 2d) pixels
constexpr int HL_BK = 32;
constexpr int HL_MAX_ACT_STAGES = 4, HL_W_STAGES = 4;   // activation stages: 2, up to 4 for layers with resident weights
constexpr int HL_CONV_THREADS = 256;
constexpr int HL_THREADS = 64 + 128 + HL_CONV_THREADS + 32;   // act TMA, MMA, 4 epilogue, 8 converter, weight producer
constexpr float HL_SCALE = 2048.f, HL_INV_SCALE = 1.f / 2048.f;

struct HaloParams {
    const float* bias; float* y; const uint8_t* w;
    const float* mask; const float* res;
    int res_cs;
    int y_cs, mask_cs, B, H, W, Cin, Cout, cout_valid;
    int tiles_x, total_tiles, kchunks;
    int b_bytes;          // Cout * 64: one fp16 weight tile (h or l) of a (tap, slice)
    int w_stage_bytes;    // 2 * b_bytes rounded up to 1024
    int accumulate, desc_mode;
    int exp_skip_conv;    // experiment (PWC_HALO_EXP=1): converters do nothing -> wrong results, upper bound of a split-input variant
    int dil, bw;          // dilation d; box width in pixels (128 + 2d, or W + 2 in flat mode)
    int flat, nr;         // flat mode (W < 128, d = 1): a tile is 128 consecutive SLOTS of the padded row-major space
                          // (rows of bw = W + 2 slots); nr = box rows.  Tap (ky,kx) is still one uniform shift ky*bw + kx.
    int tiles_img;        // tiles per image
    int row_loads;        // 1: one box with row traversal stride d (d <= 8); 3: one single-row box per row (d > 8)
    int act_stage;        // bytes per activation stage (3 * bw * 128 rounded up to 1024)
    int act_stages;       // 2..4
    int n_sets;           // independent (main | corr) accumulator sets per tile that the taps rotate over: consecutive MMAs
                          // into ONE accumulator serialise on its latency (~125 clk measured), which dominates narrow layers
    int w_resident;       // all 9 * kchunks weight images stay in shared memory (small layers): loaded once per CTA
    float alpha, mask_alpha;
    unsigned long long* dbg;   // optional timeline (clock64): 8 events x 8 tiles per CTA; nullptr in production
    // split activations (round 2): conv -> conv chains keep their intermediate tensors as [h | l * 2^11] fp16 rows (per
    // pixel and 32-channel slice 128 bytes: exactly what the converter warps produce; the cost volume's operands use the same
    // row layout with an unscaled l),
    // written by the producer's epilogue and landed by TMA ready for the MMAs: the consumer's converter pass disappears.
    int in_split;              // the input tensor map is over split rows (fp16): the converter warps only forward the barrier
    __half* ys; int ys_cs;     // optional split output (halfs per pixel = 2 * channels of that tensor); y may then be null
};

#define HL_DBG(ev, tile) do { if (dbg && (tile) < 8) dbg[(ev) * 8 + (tile)] = clock64(); } while (0)

__device__ __forceinline__ void hl_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__global__ void __launch_bounds__(HL_THREADS, 1)
conv3x3_tc_halo_kernel(const __grid_constant__ CUtensorMap tmX, const HaloParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    // barriers: act_full[2], act_conv[2], act_empty[2], w_full[4], w_empty[4], acc_full[2], acc_empty[2]
    __shared__ __align__(8) uint64_t bars[3 * HL_MAX_ACT_STAGES + 2 * HL_W_STAGES + 4];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int MS = HL_MAX_ACT_STAGES;
    const uint32_t bar_afull = smem_u32(&bars[0]), bar_aconv = smem_u32(&bars[MS]), bar_aempty = smem_u32(&bars[2 * MS]);
    const uint32_t bar_wfull = smem_u32(&bars[3 * MS]), bar_wempty = smem_u32(&bars[3 * MS + HL_W_STAGES]);
    const uint32_t bar_accf = smem_u32(&bars[3 * MS + 2 * HL_W_STAGES]), bar_acce = smem_u32(&bars[3 * MS + 2 * HL_W_STAGES + 2]);
    const int AS = p.act_stages;
    const uint32_t w_base = base + p.act_stages * p.act_stage;
    const int n_rows = (p.flat ? p.nr : HL_BH) * p.bw;

    if (threadIdx.x == 0) {
        for (int s = 0; s < HL_MAX_ACT_STAGES; ++s) {
            mbar_init(bar_afull + 8 * s, 1);
            mbar_init(bar_aconv + 8 * s, HL_CONV_THREADS);
            mbar_init(bar_aempty + 8 * s, 1);
        }
        for (int s = 0; s < HL_W_STAGES; ++s) {
            mbar_init(bar_wfull + 8 * s, 1);
            mbar_init(bar_wempty + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_accf + 8 * a, 1);
            mbar_init(bar_acce + 8 * a, 4);
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
    const int tiles_per_img = p.tiles_img;
    unsigned long long* dbg = p.dbg ? p.dbg + (size_t)blockIdx.x * 64 : nullptr;

    if (warp == 0) {
        // ===================== activation producer =====================
        if (elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
            int it = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const int b = t / tiles_per_img, r = t - b * tiles_per_img;
                int y, x0;
                if (p.flat) { y = (r * HL_M) / p.bw; x0 = 0; }           // first row touched by the tile's slots
                else { y = r / p.tiles_x; x0 = (r - y * p.tiles_x) * HL_M; }
                for (int c = 0; c < KC; ++c, ++it) {
                    const int s = it % AS;
                    mbar_wait(bar_aempty + 8 * s, ((it / AS) & 1) ^ 1);
                    HL_DBG(0, it);
                    mbar_expect_tx(bar_afull + 8 * s, (uint32_t)n_rows * 128);
                    const int c0 = c * (p.in_split ? 2 * HL_BK : HL_BK);      // element offset of the slice (fp32 or halfs)
                    if (p.row_loads == 1) {
                        tma_load_4d(base + s * p.act_stage, &tmX, bar_afull + 8 * s, c0, x0 - p.dil, y - p.dil, b);
                    } else {
#pragma unroll
                        for (int r2 = 0; r2 < HL_BH; ++r2)
                            tma_load_4d(base + s * p.act_stage + r2 * p.bw * 128, &tmX, bar_afull + 8 * s, c0, x0 - p.dil,
                                        y + (r2 - 1) * p.dil, b);
                    }
                }
            }
        }
    } else if (warp == 14) {
        // ===================== weight producer: [W_h | W_l] image of every (slice, tap), one bulk copy each =====================
        if (elect_one()) {
            int wt = 0;
            const uint32_t bytes = 2 * p.b_bytes;
            if (p.w_resident) {
                // small layers: the weight stream would be latency-bound (a tap's MMAs are shorter than an L2 round trip),
                // so every (slice, tap) image is loaded ONCE per persistent CTA
                mbar_expect_tx(bar_wfull, (uint32_t)(9 * KC) * bytes);
                for (int c = 0; c < KC; ++c)
                    for (int tap = 0; tap < 9; ++tap)
                        bulk_load_1d(w_base + (c * 9 + tap) * p.w_stage_bytes, p.w + (size_t)(tap * KC + c) * bytes, bytes, bar_wfull);
            } else {
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                for (int c = 0; c < KC; ++c) {
                    for (int tap = 0; tap < 9; ++tap, ++wt) {
                        const int s = wt & (HL_W_STAGES - 1);
                        mbar_wait(bar_wempty + 8 * s, ((wt / HL_W_STAGES) & 1) ^ 1);
                        mbar_expect_tx(bar_wfull + 8 * s, bytes);
                        bulk_load_1d(w_base + s * p.w_stage_bytes, p.w + (size_t)(tap * KC + c) * bytes, bytes, bar_wfull + 8 * s);
                    }
                }
            }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            const bool leader = true;
            const uint32_t tmem_u = tmem_acc, base_u = base, w_base_u = w_base;
            const uint32_t idesc_n = (1u << 4) | ((uint32_t)(p.Cout >> 3) << 17) | ((uint32_t)(HL_M >> 4) << 24);
            const uint32_t idesc_w = (1u << 4) | ((uint32_t)((2 * p.Cout) >> 3) << 17) | ((uint32_t)(HL_M >> 4) << 24);
            // A: K-major rows of 128 bytes ([h | l]), 128B swizzle, 8-row groups 1024 bytes apart
            const uint64_t adesc_hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
            // B: K-major rows of 64 bytes, 64B swizzle, 8-row groups 512 bytes apart
            const uint64_t bdesc_hi = ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
            // Everything that does not change is hoisted out of the issue loop (the nine tap offsets of the A descriptor,
            // the accumulator-set stride, the weight-image stride): on narrow layers the issuing thread is the critical path.
            uint32_t tapoff[9];
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) tapoff[tap] = (uint32_t)((tap / 3) * p.bw + (tap % 3) * p.dil) * 8;   // 128-byte rows, >> 4
            const uint32_t setmask = (uint32_t)p.n_sets - 1, cout2 = 2 * p.Cout, n_sets = p.n_sets;
            const uint32_t wsb16 = (uint32_t)p.w_stage_bytes >> 4;
            const bool resident = p.w_resident != 0;
            int it = 0, wt = 0, tcount = 0;
            if (resident) { mbar_wait(bar_wfull, 0); tc_fence_after(); }
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++tcount) {
                const int a = tcount & 1, u = tcount >> 1;
                if (u > 0) {                       // the epilogue has drained this accumulator set
                    mbar_wait(bar_acce + 8 * a, (u - 1) & 1);
                    tc_fence_after();
                }
                const uint32_t d_tile = tmem_u + a * 256;
                uint32_t c0off = 0;                                   // flat mode: the tile starts c0 slots into its first row
                if (p.flat) { const int r = t % tiles_per_img; c0off = (uint32_t)((r * HL_M) % p.bw) * 8; }
                for (int c = 0; c < KC; ++c, ++it) {
                    const int s = it % AS;
                    mbar_wait(bar_aconv + 8 * s, (it / AS) & 1);
                    if (leader) HL_DBG(3, it);
                    tc_fence_after();
                    const uint32_t ast = base_u + s * p.act_stage;
                    const bool ks2 = p.Cin - c * HL_BK > 16;             // channels 16..31 of the slice are zero padding otherwise
                    const uint64_t ad0 = (adesc_hi | (uint64_t)(((ast >> 4) & 0x3FFF) | (1u << 16))) + c0off;
                    uint64_t bd = bdesc_hi | (uint64_t)((((w_base_u + (resident ? c * 9 * p.w_stage_bytes : 0)) >> 4) & 0x3FFF) | (1u << 16));
                    const uint32_t use0 = (uint32_t)c * 9;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        if (!resident) {
                            const int ws = wt & (HL_W_STAGES - 1);
                            mbar_wait(bar_wfull + 8 * ws, (wt / HL_W_STAGES) & 1);
                            tc_fence_after();
                            bd = bdesc_hi | (uint64_t)((((w_base_u + ws * p.w_stage_bytes) >> 4) & 0x3FFF) | (1u << 16));
                        }
                        const uint64_t ad = ad0 + tapoff[tap];
                        const uint32_t use = use0 + tap;
                        const uint32_t d_main = d_tile + (use & setmask) * cout2, d_corr = d_main + p.Cout;
                        if (leader) hl_mma(d_main, ad, bd, idesc_w, use < n_sets ? 0u : 1u);          // A_h x [W_h | W_l] -> main | corr
                        if (ks2 && leader) hl_mma(d_main, ad + 2, bd + 2, idesc_w, 1u);
                        if (leader) hl_mma(d_corr, ad + 4, bd, idesc_n, 1u);                            // A_l x W_h -> corr
                        if (ks2 && leader) hl_mma(d_corr, ad + 6, bd + 2, idesc_n, 1u);
                        if (resident) bd += wsb16;
                        else { if (leader) tc_commit(bar_wempty + 8 * (wt & (HL_W_STAGES - 1))); ++wt; }
                    }
                    if (leader) tc_commit(bar_aempty + 8 * s);
                    if (leader) HL_DBG(4, it);
                }
                if (leader) tc_commit(bar_accf + 8 * a);
            }
        }
    } else if (warp < 6) {
        // ===================== epilogue (warps 2..5; TMEM lane quadrant = warp % 4) =====================
        const int q = warp & 3;
        const int m = q * 32 + lane;
        int tcount = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++tcount) {
            const int b = t / tiles_per_img, r = t - b * tiles_per_img;
            int y, x;
            if (p.flat) { const int slot = r * HL_M + m; y = slot / p.bw; x = slot - y * p.bw; }
            else { y = r / p.tiles_x; x = (r - y * p.tiles_x) * HL_M + m; }
            const int a = tcount & 1, u = tcount >> 1;
            mbar_wait(bar_accf + 8 * a, u & 1);
            if (threadIdx.x == 64) HL_DBG(5, tcount);
            tc_fence_after();
            const bool valid = x < p.W && y < p.H;
            const size_t pix = ((size_t)b * p.H + y) * p.W + x;
            float* yrow = p.y + pix * p.y_cs;
            const float* mrow = p.mask ? p.mask + pix * p.mask_cs : nullptr;
            const float* rrow = p.res ? p.res + pix * p.res_cs : nullptr;
            const bool vec = ((p.y_cs & 3) == 0) && aligned16(p.y) && ((p.cout_valid & 3) == 0) && !p.res &&
                             (!p.mask || (((p.mask_cs & 3) == 0) && aligned16(p.mask)));
            const uint32_t tbase = tmem_acc + ((uint32_t)(q * 32) << 16) + a * 256;
            for (int n0 = 0; n0 < p.Cout; n0 += 16) {
                uint32_t rm[16], rc[16];
                float acc[16], accc[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) { acc[j] = 0.f; accc[j] = 0.f; }
                for (int st = 0; st < p.n_sets; ++st) {
                    tmem_ld16(tbase + st * 2 * p.Cout + p.Cout + n0, rc);
                    tmem_ld16(tbase + st * 2 * p.Cout + n0, rm);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) { acc[j] += __uint_as_float(rm[j]); accc[j] += __uint_as_float(rc[j]); }
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] += accc[j] * HL_INV_SCALE;
                if (valid) {
                    if (p.bias) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) acc[j] += __ldg(p.bias + n0 + j);
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[j] = leaky(acc[j], p.alpha);
                    if (p.ys) {
                        // 16 channels of this pixel as split rows: h at half (n0 / 32) * 64 + n0 % 32 of the pixel, l 32 halfs later
                        __half* hp = p.ys + pix * p.ys_cs + (n0 >> 5) * 64 + (n0 & 31);
                        uint32_t hw[8], lw[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const __half2 h2 = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
                            const float2 f2 = __half22float2(h2);
                            const __half2 l2 = __floats2half2_rn((acc[2 * j] - f2.x) * HL_SCALE, (acc[2 * j + 1] - f2.y) * HL_SCALE);
                            hw[j] = *reinterpret_cast<const uint32_t*>(&h2);
                            lw[j] = *reinterpret_cast<const uint32_t*>(&l2);
                        }
                        uint4* hq = reinterpret_cast<uint4*>(hp);
                        uint4* lq = reinterpret_cast<uint4*>(hp + 32);
                        hq[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]); hq[1] = make_uint4(hw[4], hw[5], hw[6], hw[7]);
                        lq[0] = make_uint4(lw[0], lw[1], lw[2], lw[3]); lq[1] = make_uint4(lw[4], lw[5], lw[6], lw[7]);
                    }
                    if (!p.y) {
                    } else if (vec) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            if (n0 + j >= p.cout_valid) break;
                            float4 v = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
                            if (mrow) {
                                const float4 mk = ldg4(mrow + n0 + j);
                                v.x *= mk.x > 0.f ? 1.f : p.mask_alpha; v.y *= mk.y > 0.f ? 1.f : p.mask_alpha;
                                v.z *= mk.z > 0.f ? 1.f : p.mask_alpha; v.w *= mk.w > 0.f ? 1.f : p.mask_alpha;
                            }
                            float4* dst = reinterpret_cast<float4*>(yrow + n0 + j);
                            if (p.accumulate) { const float4 o = *dst; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                            *dst = v;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            if (n0 + j >= p.cout_valid) break;
                            float v = acc[j];
                            if (mrow) v *= __ldg(mrow + n0 + j) > 0.f ? 1.f : p.mask_alpha;
                            if (rrow) v += __ldg(rrow + n0 + j);
                            if (p.accumulate) v += yrow[n0 + j];
                            yrow[n0 + j] = v;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acce + 8 * a);
            if (threadIdx.x == 64) HL_DBG(6, tcount);
        }
    } else {
        // ===================== converters (warps 6..13): fp32 pixel row -> [h | l * 2^11] fp16, in place =====================
        const int ct = threadIdx.x - 192;   // 0..255
        int it = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            for (int c = 0; c < KC; ++c, ++it) {
                const int s = it % AS;
                mbar_wait(bar_afull + 8 * s, (it / AS) & 1);
                if (ct == 0) HL_DBG(1, it);
                uint8_t* stp = base_ptr + (size_t)s * p.act_stage;
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int R = ct + rr * HL_CONV_THREADS;
                    if (R < n_rows && !p.exp_skip_conv && !p.in_split) {
                        uint8_t* row = stp + (size_t)R * 128;
                        const int sw = R & 7;              // 128B swizzle: logical 16-byte chunk j sits at chunk j ^ (R & 7)
                        float4 v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = *reinterpret_cast<const float4*>(row + ((j ^ sw) << 4));
                        uint4 hq[4], lq[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 a = v[2 * j], bq = v[2 * j + 1];
                            const __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
                            const __half2 h2 = __floats2half2_rn(bq.x, bq.y), h3 = __floats2half2_rn(bq.z, bq.w);
                            const float2 f0 = __half22float2(h0), f1 = __half22float2(h1), f2 = __half22float2(h2), f3 = __half22float2(h3);
                            const __half2 l0 = __floats2half2_rn((a.x - f0.x) * HL_SCALE, (a.y - f0.y) * HL_SCALE);
                            const __half2 l1 = __floats2half2_rn((a.z - f1.x) * HL_SCALE, (a.w - f1.y) * HL_SCALE);
                            const __half2 l2 = __floats2half2_rn((bq.x - f2.x) * HL_SCALE, (bq.y - f2.y) * HL_SCALE);
                            const __half2 l3 = __floats2half2_rn((bq.z - f3.x) * HL_SCALE, (bq.w - f3.y) * HL_SCALE);
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
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                if (ct == 0) HL_DBG(2, it);
                mbar_arrive(bar_aconv + 8 * s);
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(512));
    }
}

// Returns CONV_HALO_UNSUPPORTED when the arguments need the streaming kernel of conv_tc_f16.cu.
int launch_conv_halo(const float* x, int x_cs, const void* w_packed, const float* bias, float* y, int y_cs,
                     int B, int H, int W, int Cin, int Cout, int dilation, float alpha, const float* mask, int mask_cs,
                     float mask_alpha, int accumulate, int cout_valid, const float* res, int res_cs, cudaStream_t st,
                     int in_split, void* y_split, int ys_cs) {
    EncodeTiledFn enc = get_encode();
    if (in_split && (Cin & 31)) return -1000;                   // split rows come in whole 32-channel slices
    if (y_split && ((Cout & 31) || (ys_cs & 15) || !aligned16(y_split))) return -1000;
    if (!enc || Cout > 128 || (Cout & 7) || dilation < 1 || dilation > 16) return -1000;   // two accumulator sets of 2*Cout columns must fit 512
    // rows narrower than a tile: flat mode (d = 1 only) packs several rows into the 128 MMA rows
    const int flat = (W < HL_M && dilation == 1 && !getenv("PWC_HALO_NO_FLAT")) ? 1 : 0;
    const int nr = flat ? (HL_M - 1 + (W + 2) - 1) / (W + 2) + 3 : 0;
    CUtensorMap tmX;
    {
        // x_cs counts elements of the input tensor: floats, or halfs of a split tensor (2 * its channel count)
        const cuuint64_t esz = in_split ? 2 : 4;
        cuuint64_t dims[4] = {(cuuint64_t)(in_split ? 2 * Cin : Cin), (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)x_cs * esz, (cuuint64_t)W * x_cs * esz, (cuuint64_t)H * W * x_cs * esz};
        // rows y-d, y, y+d through the traversal stride of the row dimension (TMA allows strides up to 8); larger
        // dilations use one single-row box per row (their row pitch (128+2d)*128 B must keep the 1024-byte swizzle phase)
        const bool strided = dilation <= 8;
        if (!strided && (((HL_M + 2 * dilation) * 128) & 1023)) return -1000;
        cuuint32_t box[4] = {(cuuint32_t)(in_split ? 2 * HL_BK : HL_BK), (cuuint32_t)(HL_M + 2 * dilation), (cuuint32_t)(strided ? HL_BH * dilation : 1), 1};
        cuuint32_t es[4] = {1, 1, (cuuint32_t)(strided ? dilation : 1), 1};
        if (flat) { box[1] = (cuuint32_t)(W + 2); box[2] = (cuuint32_t)nr; es[2] = 1; }
        CUresult r = enc(&tmX, in_split ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)x, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv3x3_tc_halo: cuTensorMapEncodeTiled(x) failed with %d", (int)r); return PWC_E_BADARG; }
    }
    HaloParams p{};
    p.bias = bias; p.y = y; p.w = (const uint8_t*)w_packed; p.mask = mask; p.res = res; p.res_cs = res_cs;
    p.y_cs = y_cs; p.mask_cs = mask_cs; p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.cout_valid = cout_valid;
    p.tiles_x = (W + HL_M - 1) / HL_M;
    p.flat = flat; p.nr = nr;
    p.tiles_img = flat ? (H * (W + 2) + HL_M - 1) / HL_M : p.tiles_x * H;
    const long long tiles = (long long)p.tiles_img * B;
    if (tiles >= (1ll << 30)) return -1000;
    p.total_tiles = (int)tiles;
    p.kchunks = (Cin + HL_BK - 1) / HL_BK;
    p.b_bytes = Cout * 64;
    p.w_stage_bytes = (2 * p.b_bytes + 1023) / 1024 * 1024;
    p.accumulate = accumulate; p.alpha = alpha; p.mask_alpha = mask_alpha;
    p.in_split = in_split; p.ys = (__half*)y_split; p.ys_cs = ys_cs;
    p.dil = dilation; p.bw = flat ? W + 2 : HL_M + 2 * dilation; p.row_loads = dilation <= 8 ? 1 : HL_BH;
    p.act_stage = ((flat ? nr : HL_BH) * p.bw * 128 + 1023) / 1024 * 1024;
    p.desc_mode = 0;
    p.exp_skip_conv = getenv("PWC_HALO_EXP") ? 1 : 0;
    if (const char* e = getenv("PWC_HALO_DESC")) p.desc_mode = atoi(e);
    // Rotating accumulator sets (power of two, <= 8, n_sets * 2 * Cout <= 256 columns per tile) paid off while the MMA
    // issue was slowed by the lane-0 wrapper; with elect.sync one set is fastest (the epilogue reads every set back:
    // 2.4k clk per 16-channel tile with 8 sets), so 1 is the default and PWC_HALO_SETS selects more.
    int max_sets = 1;
    while (max_sets < 8 && 2 * max_sets * 2 * Cout <= 256) max_sets *= 2;
    p.n_sets = 1;
    if (const char* e = getenv("PWC_HALO_SETS")) { int v = atoi(e); if ((v == 2 || v == 4 || v == 8) && v <= max_sets) p.n_sets = v; }
    p.act_stages = 2;   // measured: a third activation stage does not help (174.6 vs 171.5 us at 128->128), the limit is operand bandwidth
    int want_stages = 0;
    if (const char* e = getenv("PWC_HALO_STAGES")) want_stages = atoi(e);
    if (want_stages == 3) p.act_stages = 3;
    const size_t w_all = (size_t)9 * p.kchunks * p.w_stage_bytes;
    p.w_resident = (2 * (size_t)p.act_stage + w_all + 1024 <= 227 * 1024) && !getenv("PWC_HALO_NO_RESIDENT");
    if (p.w_resident) {
        // small layers (resident weights): a stage is held from the TMA issue to the last MMA that reads it (~5k clk
        // at 16->16: 1.9k TMA latency + 1.4k conversion + 1.7k MMAs), so two stages cap the tile period at ~2.5k clk
        p.act_stages = 2;
        const int fit = (int)((227 * 1024 - 1024 - w_all) / p.act_stage);
        const int lim = want_stages >= 2 && want_stages <= HL_MAX_ACT_STAGES ? want_stages : HL_MAX_ACT_STAGES;
        if (fit > 2) p.act_stages = fit < lim ? fit : lim;
    }
    size_t smem = (size_t)p.act_stages * p.act_stage + (p.w_resident ? w_all : (size_t)HL_W_STAGES * p.w_stage_bytes) + 1024;
    if (smem > 227 * 1024) { p.act_stages = 2; smem = (size_t)2 * p.act_stage + (size_t)HL_W_STAGES * p.w_stage_bytes + 1024; }
    if (smem > 227 * 1024) return -1000;
    cudaError_t e = cudaFuncSetAttribute(conv3x3_tc_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("conv3x3_tc_halo: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    const int grid = p.total_tiles < 148 ? p.total_tiles : 148;
    static unsigned long long* dbg_buf = nullptr;
    if (getenv("PWC_HALO_DEBUG")) {
        if (!dbg_buf) cudaMalloc(&dbg_buf, 148 * 64 * 8);
        cudaMemsetAsync(dbg_buf, 0, 148 * 64 * 8, st);
        p.dbg = dbg_buf;
    }
    conv3x3_tc_halo_kernel<<<grid, HL_THREADS, smem, st>>>(tmX, p);
    PWC_CHECK_LAUNCH("conv3x3_tc_halo_kernel");
    if (p.dbg) {   // debugging aid only (synchronises): timeline of the first chunks / tiles of one CTA
        cudaStreamSynchronize(st);
        static int printed = 0;
        if (printed++ < 1) {
            unsigned long long h[64];
            const char* names[8] = {"act_tma", "full_seen", "conv_done", "mma_start", "mma_issued", "acc_seen", "epi_done", "-"};
            cudaMemcpy(h, p.dbg + 64 * (grid / 2), 64 * 8, cudaMemcpyDeviceToHost);
            fprintf(stderr, "[halo dbg] cta %d Cin %d Cout %d kchunks %d resident %d (clk from first TMA issue; columns = chunks / tiles 0..7)\n", grid / 2, Cin, Cout, p.kchunks, p.w_resident);
            for (int e = 0; e < 7; ++e) {
                fprintf(stderr, "   %-10s", names[e]);
                for (int t = 0; t < 8; ++t) fprintf(stderr, " %7lld", (long long)(h[e * 8 + t] - h[0]));
                fprintf(stderr, "\n");
            }
        }
    }
    return 0;
}

}  // namespace pwc
