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
//
// Round 2 (what the per-CTA timelines of PWC_HALO_DEBUG=1 led to; DESIGN.md 3.2, profiles/r02_halo_epilogue.log):
//   * split activations: input and/or output as [h | l * 2^11] fp16 rows (conv -> conv chains skip the converter pass);
//   * epilogue: bias in shared memory, 32-channel passes, per-warp swizzled staging + TMA tensor stores (dgrad: the leaky
//     mask tile arrives by TMA, accumulate is a reduce-add store), accumulators released right after tcgen05.wait::ld,
//     software-pipelined for the 16- / 32-channel layers (the next tile's tcgen05.ld in flight during the stores);
//   * MMA issue loop with 32-bit descriptor arithmetic (3.6 instructions per MMA), no integer division in any role;
//   * work items: strided whole tiles, then a channel-split tail round; 16-deep weight ring, per-slice residency barriers;
//   * stride 2 as a 2 x 2 convolution over a space-to-depth view read through a 5-D tensor map (s2d);
//   * 16-channel inputs: 64-byte shared-memory rows (64B swizzle) instead of half-empty 128-byte ones;
//   * programmatic dependent launch: the prologue overlaps the previous kernel's tail (common.cuh).
#include "tc_common.cuh"
#include <cuda_fp16.h>
#include <cstdlib>
#include <cstring>

namespace pwc {

constexpr int HL_M = 128;                          // output pixels per tile (one row segment)
constexpr int HL_BH = 3;                           // halo box: 3 rows (y-d, y, y+d) x (128 + 2d) pixels
constexpr int HL_BK = 32;
constexpr int HL_MAX_ACT_STAGES = 8;                   // activation stages: 2 (streamed weights), up to 8 for layers with resident weights
constexpr int HL_W_STAGES = 16;                        // barriers of the weight ring; the ring itself is p.w_stages (4..16, power of two) deep
constexpr int HL_W_MIN_STAGES = 4;
constexpr int HL_CONV_THREADS = 256;
constexpr int HL_THREADS = 64 + 128 + HL_CONV_THREADS + 32;   // act TMA, MMA, 4 epilogue, 8 converter, weight producer
constexpr size_t HL_SMEM_BUDGET = 225 * 1024;         // dynamic shared memory: 227 KB per CTA minus the static barriers + bias (~1.2 KB)
constexpr float HL_SCALE = 2048.f, HL_INV_SCALE = 1.f / 2048.f;

struct HaloParams {
    const float* bias; float* y; const uint8_t* w;
    const float* mask; const float* res;
    int res_cs;
    int y_cs, mask_cs, B, H, W, Cin, Cout, cout_valid;
    int tiles_x, total_tiles, kchunks;
    // Work items: CTA i processes the whole tiles i, i + grid, ... (`rounds` of them), then at most one item of the TAIL -- the
    // last total_tiles - rounds * grid tiles, each split n_split ways along the output channels (cn_split channels per item,
    // tail_items = tail tiles * n_split <= grid) so that the partly filled last round costs 1 / n_split of a tile time (layers
    // with fewer tiles than SMs are all tail).  n_split = 1: no split possible (tail_items = tail tiles, cn_split = Cout).
    int rounds, tail_items, n_split, cn_split;
    int b_bytes;          // Cout * 64: one fp16 weight tile (h or l) of a (tap, slice)
    int w_stage_bytes;    // 2 * b_bytes rounded up to 1024
    int accumulate, desc_mode;
    int exp_skip_conv;    // experiment (PWC_HALO_EXP=1): converters do nothing -> wrong results, upper bound of a split-input variant
    // Tap set.  Stride 1: the nine taps (ky, kx) of a 3 x 3 kernel, box rows y-d, y, y+d.  Stride 2 ("space to depth", s2d = 1):
    // the input is read through a 5-D tensor map as X'[y][x][(py, px, c)] = X[2y+py][2x+px][c] (4 * Cin channels at half the
    // resolution), on which the stride-2 3 x 3 convolution is a stride-1 2 x 2 one: taps (dy, dx) in {0,1}^2, box rows y, y+1,
    // no padding before (TF 'SAME' with even sizes pads one row / column AFTER), weights re-indexed host-side with zeros
    // where 2*dy+py or 2*dx+px would be 3.  A K slice is 32 consecutive (px, c) values of one row phase py.
    int n_taps, box_rows, pad, s2d, kc_per_py;
    int tap_dy[9], tap_dx[9], tap_id[9];     // tap offsets in box rows / pixels (times dil for dx), and the tap's index in the packed weights
    int tma_y, tma_ys;    // the epilogue stores the fp32 / split output through TMA (shared-memory staging + bulk tensor store)
    int epi_off;          // byte offset of the epilogue staging area: 4 warps x 2 buffers x (32 pixels x 128 or 64 bytes)
    int row64;            // 16-channel input, row tiles: 64-byte shared-memory rows (box of 16 channels, 64B swizzle) instead of half-empty 128-byte ones
    int single_pass;      // Cout = 16 or 32, one plainly TMA-stored output (no mask, no reduce-add), no channel-split tail: the software-pipelined epilogue
    int tma_mask;         // dgrad: the leaky-derivative mask tile of a pass is TMA-loaded into a per-warp buffer behind the staging area
    int tma_red;          // accumulate through the TMA engine's reduce-add store (UTMAREDG) instead of a read-modify-write
    int exp_direct_store; // experiment (PWC_HALO_EXP=3): 16-byte-per-lane stores (round-1 pattern)
    int exp_skip_store;   // experiment (PWC_HALO_EXP=2): the epilogue stores nothing -> upper bound of the store path
    int dil, bw;          // dilation d; box width in pixels (128 + 2d, or W + 2 in flat mode)
    int flat, nr;         // flat mode (W < 128, d = 1): a tile is 128 consecutive SLOTS of the padded row-major space
                          // (rows of bw = W + 2 slots); nr = box rows.  Tap (ky,kx) is still one uniform shift ky*bw + kx.
    int tiles_img;        // tiles per image
    int row_loads;        // 1: one box with row traversal stride d (d <= 8); 3: one single-row box per row (d > 8)
    int act_stage;        // bytes per activation stage (3 * bw * 128 rounded up to 1024)
    int act_stages;       // 2..4
    int w_stages, w_shift; // depth of the weight ring (streaming layers) and its log2
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

// tcgen05.mma with the shared-memory descriptors given as (lo, hi) words: only the low word (start address) changes
// between the MMAs of a tile, so the issuing thread does 32-bit adds instead of 64-bit descriptor arithmetic.
constexpr uint32_t HL_ADESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);   // A: 128-byte K-major rows, 128B swizzle, 8-row groups 1024 B apart
constexpr uint32_t HL_BDESC_HI = (512u >> 4) | (1u << 14) | (4u << 29);    // B: 64-byte K-major rows, 64B swizzle, 8-row groups 512 B apart
template <bool ACC, bool ROW64 = false>
__device__ __forceinline__ void hl_mma_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    // ROW64: the A tile has 64-byte rows ([h: 16 x fp16 | l: 16 x fp16] of a 16-channel layer) in the 64B swizzle -- the B layout
    constexpr uint32_t A_HI = ROW64 ? HL_BDESC_HI : HL_ADESC_HI;
    if (ACC) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            ".reg .b64 da, db;\n\t"
            "setp.eq.u32 p, 1, 1;\n\t"
            "mov.b64 da, {%1, %2};\n\t"
            "mov.b64 db, {%3, %4};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
            "}" ::"r"(d_tmem), "r"(a_lo), "r"(A_HI), "r"(b_lo), "r"(HL_BDESC_HI), "r"(idesc) : "memory");
    } else {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            ".reg .b64 da, db;\n\t"
            "setp.ne.b32 p, %6, 0;\n\t"
            "mov.b64 da, {%1, %2};\n\t"
            "mov.b64 db, {%3, %4};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
            "}" ::"r"(d_tmem), "r"(a_lo), "r"(A_HI), "r"(b_lo), "r"(HL_BDESC_HI), "r"(idesc), "r"(accumulate) : "memory");
    }
}

struct IssueCtx {
    uint32_t idesc_n, idesc_w, tapoff[9], wsb16, cout, bar_wfull, bar_wempty, w_lo0, w_mask, w_shift;
};

// The MMAs of one 32-channel slice of a tile: nine taps x (A_h x [W_h|W_l] -> main|corr, A_l x W_h -> corr) x one or two
// K = 16 steps.  The issuing thread is the critical path of the narrow layers (a 16 -> 16 tile is 18 tiny MMAs: the
// round-1 loop spent ~55 instructions per tap, ~90 clk per MMA -- as long as a 128 x 256 x 16 MMA executes, so it also
// held the wide layers below the tensor pipe's rate; profiles/r02_halo_epilogue.log), so everything per tap is a 32-bit add.
template <bool RESIDENT, bool KS2, int NT, bool ROW64 = false>
__device__ __forceinline__ void hl_issue_chunk(const IssueCtx& cx, uint32_t d_main, uint32_t a_lo, uint32_t b_lo, uint32_t first, uint32_t& wt) {
    const uint32_t d_corr = d_main + cx.cout;
#pragma unroll
    for (int tap = 0; tap < NT; ++tap) {
        if (!RESIDENT) {
            const uint32_t ws = wt & cx.w_mask;
            mbar_wait(cx.bar_wfull + 8 * ws, (wt >> cx.w_shift) & 1);
            tc_fence_after();
            b_lo = cx.w_lo0 + ws * cx.wsb16;
        }
        const uint32_t al = a_lo + cx.tapoff[tap];
        if (tap == 0) hl_mma_lo<false, ROW64>(d_main, al, b_lo, cx.idesc_w, first);      // A_h x [W_h | W_l] -> main | corr
        else hl_mma_lo<true, ROW64>(d_main, al, b_lo, cx.idesc_w, 1u);
        if (KS2) hl_mma_lo<true, ROW64>(d_main, al + 2, b_lo + 2, cx.idesc_w, 1u);
        hl_mma_lo<true, ROW64>(d_corr, al + (ROW64 ? 2 : 4), b_lo, cx.idesc_n, 1u);        // A_l x W_h -> corr (l sits 32 / 64 bytes into the row)
        if (KS2) hl_mma_lo<true, ROW64>(d_corr, al + 6, b_lo + 2, cx.idesc_n, 1u);
        if (RESIDENT) b_lo += cx.wsb16;
        else { tc_commit(cx.bar_wempty + 8 * (wt & cx.w_mask)); ++wt; }
    }
}

// eight fp32 values -> 8 x fp16 h (16 bytes) and 8 x fp16 l * 2^11 (16 bytes)
__device__ __forceinline__ void hl_split8(const float4& a, const float4& bq, uint4& hq, uint4& lq) {
    const __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
    const __half2 h2 = __floats2half2_rn(bq.x, bq.y), h3 = __floats2half2_rn(bq.z, bq.w);
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1), f2 = __half22float2(h2), f3 = __half22float2(h3);
    const __half2 l0 = __floats2half2_rn((a.x - f0.x) * HL_SCALE, (a.y - f0.y) * HL_SCALE);
    const __half2 l1 = __floats2half2_rn((a.z - f1.x) * HL_SCALE, (a.w - f1.y) * HL_SCALE);
    const __half2 l2 = __floats2half2_rn((bq.x - f2.x) * HL_SCALE, (bq.y - f2.y) * HL_SCALE);
    const __half2 l3 = __floats2half2_rn((bq.z - f3.x) * HL_SCALE, (bq.w - f3.y) * HL_SCALE);
    hq = make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                    *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
    lq = make_uint4(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1),
                    *reinterpret_cast<const uint32_t*>(&l2), *reinterpret_cast<const uint32_t*>(&l3));
}

__device__ __forceinline__ void hl_tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

__device__ __forceinline__ void hl_tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// TMA-store epilogue pass (round 2).  Plain stores from the epilogue warps cost ~275 clk per STG.128 whatever their
// pattern (16 bytes per lane, 32-byte lane pairs, or fully coalesced after a shared-memory transpose: all measured,
// profiles/r02_halo_epilogue.log) -- the SM's store path accepts ~7 B/clk from these warps, and the epilogue, not the
// MMAs, set the tile period of the narrow layers.  Here a warp writes its 32 pixels x G channels into a private,
// swizzled staging buffer (conflict-free STS.128: the swizzle the tensor map undoes), and lane 0 hands the 32-pixel box
// to the TMA engine; the warp goes on with the next pass / tile while the engine drains it (two buffers per warp,
// cp.async.bulk.wait_group.read before a buffer is rewritten).  Pixels beyond the row end are clipped by the tensor map.
template <int G>
__device__ __forceinline__ void hl_stage_and_store(const CUtensorMap* map, uint8_t* buf, const uint32_t (&w)[G], int lane,
                                                   int c0, int x, int y, int b, bool reduce_add = false) {
    constexpr int NCH = G / 4;                                // 16-byte chunks per pixel row (8: 128 B, 4: 64 B)
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");    // the store that last read this buffer is done
    __syncwarp();
    const int sw = NCH == 8 ? (lane & 7) : ((lane >> 1) & 3);  // 128B / 64B swizzle of row `lane` (buffers are 1024-byte aligned)
    uint8_t* row = buf + lane * (16 * NCH);
#pragma unroll
    for (int j = 0; j < NCH; ++j)
        *reinterpret_cast<uint4*>(row + ((j ^ sw) << 4)) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        if (reduce_add)
            asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                         ::"l"(map), "r"(smem_u32(buf)), "r"(c0), "r"(x), "r"(y), "r"(b) : "memory");
        else hl_tma_store_4d(map, smem_u32(buf), c0, x, y, b);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
}

// Stores of one pass, 32 contiguous bytes per lane PAIR: after tcgen05.ld a lane holds one pixel's channels, and a plain
// 16-byte store per lane puts 32 half-written sectors per instruction on the wire (the other half follows an instruction
// later; measured ~375 clk per such instruction, profiles/r02_halo_epilogue.log).  Lanes 2k / 2k+1 swap every second
// 16-byte chunk instead: store A covers chunks (c, c+1) of the even lane's pixel, store B the same chunks of the odd
// lane's pixel -- every request is a whole 32-byte sector.  `own` / `oth` point at this lane's / the partner's pixel
// (at the pass's first channel); the mask (dgrad: leaky derivative of the forward activation) and the accumulate
// operand are read at the stored positions.  Executed by all 32 lanes (shuffles); invalid pixels only skip the stores.
template <int NCH>
__device__ __forceinline__ void hl_store_pairs(const uint32_t (&w)[4 * NCH], char* own, char* oth, bool valid, bool valid_o, bool odd,
                                               const float* mask_own, const float* mask_oth, float mask_alpha, bool accumulate) {
    char* row_e = odd ? oth : own;                     // the even lane's pixel, the odd lane's pixel
    char* row_o = odd ? own : oth;
    const bool ok_e = odd ? valid_o : valid, ok_o = odd ? valid : valid_o;
    const char* m_e = reinterpret_cast<const char*>(odd ? mask_oth : mask_own);
    const char* m_o = reinterpret_cast<const char*>(odd ? mask_own : mask_oth);
#pragma unroll
    for (int c = 0; c < NCH; c += 2) {
        uint4 mine_e, mine_o;                          // even lane keeps chunk c, sends c+1; odd lane keeps c+1, sends c
        uint4 snd;
        snd.x = odd ? w[4 * c] : w[4 * c + 4]; snd.y = odd ? w[4 * c + 1] : w[4 * c + 5];
        snd.z = odd ? w[4 * c + 2] : w[4 * c + 6]; snd.w = odd ? w[4 * c + 3] : w[4 * c + 7];
        uint4 rcv;
        rcv.x = __shfl_xor_sync(0xffffffffu, snd.x, 1); rcv.y = __shfl_xor_sync(0xffffffffu, snd.y, 1);
        rcv.z = __shfl_xor_sync(0xffffffffu, snd.z, 1); rcv.w = __shfl_xor_sync(0xffffffffu, snd.w, 1);
        const uint4 keep = odd ? make_uint4(w[4 * c + 4], w[4 * c + 5], w[4 * c + 6], w[4 * c + 7])
                               : make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
        mine_e = odd ? rcv : keep;                     // -> even pixel, chunk c + odd
        mine_o = odd ? keep : rcv;                     // -> odd pixel,  chunk c + odd
        const int off = (c + (odd ? 1 : 0)) * 16;
        if (mask_own) {
            if (ok_e) {
                const float4 mk = __ldg(reinterpret_cast<const float4*>(m_e + off));
                mine_e.x = __float_as_uint(__uint_as_float(mine_e.x) * (mk.x > 0.f ? 1.f : mask_alpha));
                mine_e.y = __float_as_uint(__uint_as_float(mine_e.y) * (mk.y > 0.f ? 1.f : mask_alpha));
                mine_e.z = __float_as_uint(__uint_as_float(mine_e.z) * (mk.z > 0.f ? 1.f : mask_alpha));
                mine_e.w = __float_as_uint(__uint_as_float(mine_e.w) * (mk.w > 0.f ? 1.f : mask_alpha));
            }
            if (ok_o) {
                const float4 mk = __ldg(reinterpret_cast<const float4*>(m_o + off));
                mine_o.x = __float_as_uint(__uint_as_float(mine_o.x) * (mk.x > 0.f ? 1.f : mask_alpha));
                mine_o.y = __float_as_uint(__uint_as_float(mine_o.y) * (mk.y > 0.f ? 1.f : mask_alpha));
                mine_o.z = __float_as_uint(__uint_as_float(mine_o.z) * (mk.z > 0.f ? 1.f : mask_alpha));
                mine_o.w = __float_as_uint(__uint_as_float(mine_o.w) * (mk.w > 0.f ? 1.f : mask_alpha));
            }
        }
        if (accumulate) {
            if (ok_e) {
                const uint4 o = *reinterpret_cast<const uint4*>(row_e + off);
                mine_e.x = __float_as_uint(__uint_as_float(mine_e.x) + __uint_as_float(o.x)); mine_e.y = __float_as_uint(__uint_as_float(mine_e.y) + __uint_as_float(o.y));
                mine_e.z = __float_as_uint(__uint_as_float(mine_e.z) + __uint_as_float(o.z)); mine_e.w = __float_as_uint(__uint_as_float(mine_e.w) + __uint_as_float(o.w));
            }
            if (ok_o) {
                const uint4 o = *reinterpret_cast<const uint4*>(row_o + off);
                mine_o.x = __float_as_uint(__uint_as_float(mine_o.x) + __uint_as_float(o.x)); mine_o.y = __float_as_uint(__uint_as_float(mine_o.y) + __uint_as_float(o.y));
                mine_o.z = __float_as_uint(__uint_as_float(mine_o.z) + __uint_as_float(o.z)); mine_o.w = __float_as_uint(__uint_as_float(mine_o.w) + __uint_as_float(o.w));
            }
        }
        if (ok_e) *reinterpret_cast<uint4*>(row_e + off) = mine_e;
        if (ok_o) *reinterpret_cast<uint4*>(row_o + off) = mine_o;
    }
}

// One epilogue pass over G channels of one pixel (a lane of the accumulator quadrant): TMEM -> registers, main + 2^-11 x
// correction, bias (from shared memory: a global load here sits on the tile's critical path at L2 latency -- measured
// ~1k clk per pass, profiles/r02_halo_epilogue.log), leaky, stores.  G = 32 when the layer has whole 32-channel groups:
// half the TMEM round trips (~0.5-1k clk each under MMA load) and a lane writes whole 128-byte lines (its pixel's fp32
// channels, or the [h | l] split row of the slice).  A shared-memory transpose for fully coalesced stores was measured
// SLOWER (same log): shared memory is this kernel's saturated resource.
template <int G>
__device__ __forceinline__ void hl_epilogue_pass(const HaloParams& p, const float* s_bias, uint32_t tbase, int n0, bool valid,
                                                 size_t pix, bool valid_o, size_t pix_o, bool odd, bool vec,
                                                 unsigned long long* dbg, int tcount,
                                                 const CUtensorMap* tmY, const CUtensorMap* tmYS, uint8_t* stage, int& nbuf,
                                                 int lane, int wx, int wy, int wb, uint32_t release_bar, int ch0, int cn,
                                                 const CUtensorMap* tmM, uint8_t* mask_buf, uint32_t bar_mask, uint32_t& mask_phase) {
    float acc[G];
    {
        uint32_t rm[G], rc[G];
#pragma unroll
        for (int g = 0; g < G; g += 16) {
            tmem_ld16(tbase + cn + n0 + g, rc + g);
            tmem_ld16(tbase + n0 + g, rm + g);
        }
        tmem_ld_wait();
        if (release_bar) {
            // last pass of the tile: the accumulators are in registers, hand the TMEM set back to the MMA issuer now
            // (not after the stores: the narrow layers' tile period is the epilogue's latency chain)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(release_bar);
        }
#pragma unroll
        for (int j = 0; j < G; ++j) acc[j] = __uint_as_float(rm[j]) + __uint_as_float(rc[j]) * HL_INV_SCALE;
    }
    if (n0 == 0) HL_DBG(7, tcount);
    if (p.exp_skip_store) return;
#pragma unroll
    for (int j = 0; j < G; j += 4) {
        const float4 bq = *reinterpret_cast<const float4*>(s_bias + ch0 + n0 + j);
        acc[j] = leaky(acc[j] + bq.x, p.alpha); acc[j + 1] = leaky(acc[j + 1] + bq.y, p.alpha);
        acc[j + 2] = leaky(acc[j + 2] + bq.z, p.alpha); acc[j + 3] = leaky(acc[j + 3] + bq.w, p.alpha);
    }
    if (n0 == 0) HL_DBG(8, tcount);
    if (p.ys) {
        // channels as split rows: h at half (n0 / 32) * 64 + n0 % 32 of the pixel, l 32 halfs later (G = 32: one whole 128-byte row)
        uint32_t w[G];
#pragma unroll
        for (int j = 0; j < G / 2; ++j) {
            const __half2 h2 = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
            const float2 f2 = __half22float2(h2);
            const __half2 l2 = __floats2half2_rn((acc[2 * j] - f2.x) * HL_SCALE, (acc[2 * j + 1] - f2.y) * HL_SCALE);
            w[j] = *reinterpret_cast<const uint32_t*>(&h2);
            w[G / 2 + j] = *reinterpret_cast<const uint32_t*>(&l2);
        }
        if (p.tma_ys && G == 32) {
            hl_stage_and_store<G>(tmYS, stage + (nbuf & 1) * (32 * 4 * G), w, lane, ((ch0 + n0) >> 5) * 64, wx, wy, wb);
            ++nbuf;
        } else if (G == 32 && !p.exp_direct_store) {
            char* own = reinterpret_cast<char*>(p.ys + pix * p.ys_cs + ((ch0 + n0) >> 5) * 64);
            char* oth = reinterpret_cast<char*>(p.ys + pix_o * p.ys_cs + ((ch0 + n0) >> 5) * 64);
            hl_store_pairs<G / 4>(w, own, oth, valid, valid_o, odd, nullptr, nullptr, 0.f, false);
        } else if (valid) {
            __half* hp = p.ys + pix * p.ys_cs + ((ch0 + n0) >> 5) * 64 + ((ch0 + n0) & 31);
            uint4* hq = reinterpret_cast<uint4*>(hp);
            uint4* lq = reinterpret_cast<uint4*>(hp + 32);
#pragma unroll
            for (int j = 0; j < G / 8; ++j) hq[j] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
#pragma unroll
            for (int j = 0; j < G / 8; ++j) lq[j] = make_uint4(w[G / 2 + 4 * j], w[G / 2 + 4 * j + 1], w[G / 2 + 4 * j + 2], w[G / 2 + 4 * j + 3]);
        }
    }
    if (p.y) {
        float* yrow = p.y + pix * p.y_cs + ch0;
        const float* mrow = p.mask ? p.mask + pix * p.mask_cs + ch0 : nullptr;
        const float* rrow = p.res ? p.res + pix * p.res_cs + ch0 : nullptr;
        if (p.tma_y) {
            if (p.tma_mask) {
                // dgrad: multiply by leaky'(forward activation); this warp's 32 pixel x G channel tile of the mask tensor was
                // TMA-loaded (same swizzle as the staging rows) while the MMAs / the previous pass ran
                mbar_wait(bar_mask, mask_phase);
                mask_phase ^= 1;
                constexpr int NCH = G / 4;
                const int sw = NCH == 8 ? (lane & 7) : ((lane >> 1) & 3);
                const uint8_t* mrow_s = mask_buf + lane * (16 * NCH);
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    const float4 mk = *reinterpret_cast<const float4*>(mrow_s + ((j ^ sw) << 4));
                    acc[4 * j] *= mk.x > 0.f ? 1.f : p.mask_alpha; acc[4 * j + 1] *= mk.y > 0.f ? 1.f : p.mask_alpha;
                    acc[4 * j + 2] *= mk.z > 0.f ? 1.f : p.mask_alpha; acc[4 * j + 3] *= mk.w > 0.f ? 1.f : p.mask_alpha;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the reads above before the next TMA write
                __syncwarp();
                if (lane == 0 && n0 + G < cn) {                                   // next pass of this tile
                    mbar_expect_tx(bar_mask, 32 * 4 * G);
                    tma_load_4d(smem_u32(mask_buf), tmM, bar_mask, ch0 + n0 + G, wx, wy, wb);
                }
            }
            uint32_t w[G];
#pragma unroll
            for (int j = 0; j < G; ++j) w[j] = __float_as_uint(acc[j]);
            hl_stage_and_store<G>(tmY, stage + (nbuf & 1) * (32 * 4 * G), w, lane, ch0 + n0, wx, wy, wb, p.tma_red != 0);
            ++nbuf;
        } else if (vec && p.cout_valid == p.Cout && (p.Cout % G) == 0 && !p.exp_direct_store) {
            uint32_t w[G];
#pragma unroll
            for (int j = 0; j < G; ++j) w[j] = __float_as_uint(acc[j]);
            hl_store_pairs<G / 4>(w, reinterpret_cast<char*>(yrow + n0), reinterpret_cast<char*>(p.y + pix_o * p.y_cs + ch0 + n0), valid, valid_o, odd,
                                  mrow ? mrow + n0 : nullptr, p.mask ? p.mask + pix_o * p.mask_cs + ch0 + n0 : nullptr, p.mask_alpha,
                                  p.accumulate != 0);
        } else if (!valid) {
        } else if (vec) {
#pragma unroll
            for (int j = 0; j < G; j += 4) {
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
            for (int j = 0; j < G; ++j) {
                if (n0 + j >= p.cout_valid) break;
                float v = acc[j];
                if (mrow) v *= __ldg(mrow + n0 + j) > 0.f ? 1.f : p.mask_alpha;
                if (rrow) v += __ldg(rrow + n0 + j);
                if (p.accumulate) v += yrow[n0 + j];
                yrow[n0 + j] = v;
            }
        }
    }
    if (n0 == 0) HL_DBG(9, tcount);
}

// Same as hl_stage_and_store with the 16-byte chunks produced on the fly (chunk(j) -> uint4): no G-word array is live while the
// next tile's tcgen05.ld is in flight (the pipelined epilogue below runs at the 128-register limit).
template <int G, typename F>
__device__ __forceinline__ void hl_stage_and_store_fn(const CUtensorMap* map, uint8_t* buf, int lane, int c0, int x, int y, int b, F chunk) {
    constexpr int NCH = G / 4;
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
    const int sw = NCH == 8 ? (lane & 7) : ((lane >> 1) & 3);
    uint8_t* row = buf + lane * (16 * NCH);
#pragma unroll
    for (int j = 0; j < NCH; ++j) *reinterpret_cast<uint4*>(row + ((j ^ sw) << 4)) = chunk(j);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        hl_tma_store_4d(map, smem_u32(buf), c0, x, y, b);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
}

// ---- software-pipelined epilogue of the 16-channel layers on the TMA-store path (pyramid level 1: 2 x 121 us of a forward).  Their tile period is the epilogue's latency chain (wait accumulator -> tcgen05.ld ~0.5 k clk
// -> bias / leaky -> stage -> TMA store), not any throughput.  When the NEXT tile's accumulators are already complete, its
// tcgen05.ld is issued before this tile's stores, so the TMEM read latency runs behind them.  tcgen05.ld writes its
// destination registers asynchronously until tcgen05.wait::ld: rm / rc are not touched between the issue and hl_ld_pin().
__device__ __forceinline__ uint32_t mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ void hl_pin16(uint32_t* r) {       // orders every later use of r[0..15] after this point
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                      "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]) :: "memory");
}
template <int G>
__device__ __forceinline__ void hl_ld_issue(uint32_t tbase, uint32_t (&rm)[16], uint32_t (&rc)[16]) {
#pragma unroll
    for (int g = 0; g < G; g += 16) {
        tmem_ld16(tbase + G + g, rc + g);
        tmem_ld16(tbase + g, rm + g);
    }
}
template <int G>
__device__ __forceinline__ void hl_single_item(const HaloParams& p, const float* s_bias, uint32_t tmem_q, int tcount, bool& pre,
                                               uint32_t (&rm)[16], uint32_t (&rc)[16], bool may_prefetch,
                                               uint32_t bar_accf, uint32_t bar_acce, const CUtensorMap* tmY, const CUtensorMap* tmYS,
                                               uint8_t* stage, int& nbuf, int lane, int wx, int wy, int wb) {
    const int a = tcount & 1, u = tcount >> 1;
    if (!pre) {
        mbar_wait(bar_accf + 8 * a, u & 1);
        tc_fence_after();
        hl_ld_issue<G>(tmem_q + a * 256, rm, rc);
    }
    tmem_ld_wait();
#pragma unroll
    for (int g = 0; g < G; g += 16) { hl_pin16(rm + g); hl_pin16(rc + g); }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_acce + 8 * a);              // this accumulator set is in registers
    float acc[G];
#pragma unroll
    for (int j = 0; j < G; ++j) acc[j] = __uint_as_float(rm[j]) + __uint_as_float(rc[j]) * HL_INV_SCALE;
    pre = false;
    if (may_prefetch) {
        const int a2 = (tcount + 1) & 1, u2 = (tcount + 1) >> 1;
        if (__any_sync(0xffffffffu, mbar_try(bar_accf + 8 * a2, u2 & 1))) {
            tc_fence_after();
            hl_ld_issue<G>(tmem_q + a2 * 256, rm, rc);
            pre = true;
        }
    }
#pragma unroll
    for (int j = 0; j < G; j += 4) {
        const float4 bq = *reinterpret_cast<const float4*>(s_bias + j);
        acc[j] = leaky(acc[j] + bq.x, p.alpha); acc[j + 1] = leaky(acc[j + 1] + bq.y, p.alpha);
        acc[j + 2] = leaky(acc[j + 2] + bq.z, p.alpha); acc[j + 3] = leaky(acc[j + 3] + bq.w, p.alpha);
    }
    if (p.tma_ys && G == 32) {
        // split rows [h: 32 x fp16 | l * 2^11: 32 x fp16]: chunk j < 4 = h of channels 8j .. 8j+7, chunk j >= 4 = their l
        hl_stage_and_store_fn<G>(tmYS, stage + (nbuf & 1) * (32 * 4 * G), lane, 0, wx, wy, wb, [&](int j) {
            const int c = 8 * (j & 3);
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const __half2 h2 = __floats2half2_rn(acc[c + 2 * k], acc[c + 2 * k + 1]);
                if (j < 4) o[k] = *reinterpret_cast<const uint32_t*>(&h2);
                else {
                    const float2 f2 = __half22float2(h2);
                    const __half2 l2 = __floats2half2_rn((acc[c + 2 * k] - f2.x) * HL_SCALE, (acc[c + 2 * k + 1] - f2.y) * HL_SCALE);
                    o[k] = *reinterpret_cast<const uint32_t*>(&l2);
                }
            }
            return make_uint4(o[0], o[1], o[2], o[3]);
        });
        ++nbuf;
    }
    if (p.tma_y) {
        hl_stage_and_store_fn<G>(tmY, stage + (nbuf & 1) * (32 * 4 * G), lane, 0, wx, wy, wb, [&](int j) {
            return make_uint4(__float_as_uint(acc[4 * j]), __float_as_uint(acc[4 * j + 1]), __float_as_uint(acc[4 * j + 2]), __float_as_uint(acc[4 * j + 3]));
        });
        ++nbuf;
    }
}

// 32-channel layers: the same pipeline at a granularity of 16 channels (a whole 32-channel tile in flight next to the one being
// stored does not fit 128 registers): half 1's tcgen05.ld runs behind half 0's arithmetic and staging, the next tile's half 0
// behind half 1's; the two halves fill one 128-byte staging row per pixel and leave as one TMA store.
__device__ __forceinline__ void hl_single_item32(const HaloParams& p, const float* s_bias, uint32_t tmem_q, int tcount, bool& pre,
                                                 uint32_t (&rm)[16], uint32_t (&rc)[16], bool may_prefetch,
                                                 uint32_t bar_accf, uint32_t bar_acce, const CUtensorMap* tmY, const CUtensorMap* tmYS,
                                                 uint8_t* stage, int& nbuf, int lane, int wx, int wy, int wb) {
    const int a = tcount & 1, u = tcount >> 1;
    const uint32_t tb = tmem_q + a * 256;
    if (!pre) {
        mbar_wait(bar_accf + 8 * a, u & 1);
        tc_fence_after();
        tmem_ld16(tb + 32, rc); tmem_ld16(tb, rm);                    // half 0: main columns 0..15, correction columns 32..47
    }
    uint8_t* buf = stage + (nbuf & 1) * 4096;
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");    // the store that last read this buffer is done
    __syncwarp();
    const int sw = lane & 7;
    uint8_t* row = buf + lane * 128;
    pre = false;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        tmem_ld_wait();
        hl_pin16(rm); hl_pin16(rc);
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(rm[j]) + __uint_as_float(rc[j]) * HL_INV_SCALE;
        if (half == 0) {
            tmem_ld16(tb + 48, rc); tmem_ld16(tb + 16, rm);            // half 1 in flight
        } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acce + 8 * a);             // the whole accumulator set has been read
            if (may_prefetch) {
                const int a2 = (tcount + 1) & 1, u2 = (tcount + 1) >> 1;
                if (__any_sync(0xffffffffu, mbar_try(bar_accf + 8 * a2, u2 & 1))) {
                    tc_fence_after();
                    const uint32_t tb2 = tmem_q + a2 * 256;
                    tmem_ld16(tb2 + 32, rc); tmem_ld16(tb2, rm);
                    pre = true;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
            const float4 bq = *reinterpret_cast<const float4*>(s_bias + 16 * half + j);
            acc[j] = leaky(acc[j] + bq.x, p.alpha); acc[j + 1] = leaky(acc[j + 1] + bq.y, p.alpha);
            acc[j + 2] = leaky(acc[j + 2] + bq.z, p.alpha); acc[j + 3] = leaky(acc[j + 3] + bq.w, p.alpha);
        }
        if (p.tma_ys) {
            // split row [h: 32 x fp16 | l * 2^11: 32 x fp16]: this half's h goes to chunks 2 half, 2 half + 1, its l to 4 + ...
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float v0 = acc[8 * c + 2 * k], v1 = acc[8 * c + 2 * k + 1];
                    const __half2 h2 = __floats2half2_rn(v0, v1);
                    const float2 f2 = __half22float2(h2);
                    const __half2 l2 = __floats2half2_rn((v0 - f2.x) * HL_SCALE, (v1 - f2.y) * HL_SCALE);
                    hw[k] = *reinterpret_cast<const uint32_t*>(&h2);
                    lw[k] = *reinterpret_cast<const uint32_t*>(&l2);
                }
                const int jc = 2 * half + c;
                *reinterpret_cast<uint4*>(row + ((jc ^ sw) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                *reinterpret_cast<uint4*>(row + (((jc + 4) ^ sw) << 4)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int jc = 4 * half + c;
                *reinterpret_cast<uint4*>(row + ((jc ^ sw) << 4)) =
                    make_uint4(__float_as_uint(acc[4 * c]), __float_as_uint(acc[4 * c + 1]), __float_as_uint(acc[4 * c + 2]), __float_as_uint(acc[4 * c + 3]));
            }
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        hl_tma_store_4d(p.tma_ys ? tmYS : tmY, smem_u32(buf), 0, wx, wy, wb);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    ++nbuf;
}

__global__ void __launch_bounds__(HL_THREADS, 1)
conv3x3_tc_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                       const __grid_constant__ CUtensorMap tmYS, const __grid_constant__ CUtensorMap tmM, const HaloParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    // barriers: act_full[2], act_conv[2], act_empty[2], w_full[4], w_empty[4], acc_full[2], acc_empty[2]
    __shared__ __align__(8) uint64_t bars[3 * HL_MAX_ACT_STAGES + 2 * HL_W_STAGES + 4 + 4];   // ... + mask tile of each epilogue warp
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) float s_bias[128 + 16];   // Cout <= 128; a 16-channel pass of a Cout % 16 == 8 layer reads 8 past the end

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int MS = HL_MAX_ACT_STAGES;
    const uint32_t bar_afull = smem_u32(&bars[0]), bar_aconv = smem_u32(&bars[MS]), bar_aempty = smem_u32(&bars[2 * MS]);
    const uint32_t bar_wfull = smem_u32(&bars[3 * MS]), bar_wempty = smem_u32(&bars[3 * MS + HL_W_STAGES]);
    const uint32_t bar_accf = smem_u32(&bars[3 * MS + 2 * HL_W_STAGES]), bar_acce = smem_u32(&bars[3 * MS + 2 * HL_W_STAGES + 2]);
    const int AS = p.act_stages;
    const uint32_t w_base = base + p.act_stages * p.act_stage;
    const int n_rows = (p.flat ? p.nr : p.box_rows) * p.bw;

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
        for (int q = 0; q < 4; ++q) mbar_init(bar_accf + 8 * (4 + q), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x >= 64 && threadIdx.x < 64 + 144) {
        const int c = threadIdx.x - 64;
        s_bias[c] = (p.bias && c < p.Cout) ? __ldg(p.bias + c) : 0.f;
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = tmem_base_slot;
    // Programmatic dependent launch: let the next kernel of the stream start its prologue (barriers, TMEM, resident weights)
    // on idle SMs now; whatever reads or writes activations in THIS kernel first waits for the previous one to complete
    // (griddepcontrol.wait below: the activation producer and the epilogue warps; weights and bias are never written by a
    // kernel that triggers early).  Both instructions are no-ops for a kernel launched without the attribute.
    pdl_trigger();
    const int KC = p.kchunks;
    const int tiles_per_img = p.tiles_img;
    // work items of this CTA (HaloParams): `rounds` whole tiles, then possibly one channel-split item of the tail
    const int n_items = p.rounds + ((int)blockIdx.x < p.tail_items ? 1 : 0);
    const int tail_q = blockIdx.x / p.n_split;
    const int tail_t = p.rounds * gridDim.x + tail_q;
    const int tail_ch0 = (blockIdx.x - tail_q * p.n_split) * p.cn_split;
    unsigned long long* dbg = p.dbg ? p.dbg + (size_t)blockIdx.x * 128 : nullptr;

    if (warp == 0) {
        // ===================== activation producer =====================
        if (elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
            pdl_wait();
            int it = 0, s = 0;
            uint32_t ph = 0;                                   // stage index and ring phase as counters: no division per slice
            for (int j = 0; j < n_items; ++j) {
                const int t = j < p.rounds ? (int)blockIdx.x + j * (int)gridDim.x : tail_t;
                const int b = t / tiles_per_img, r = t - b * tiles_per_img;
                int y, x0;
                if (p.flat) { y = (r * HL_M) / p.bw; x0 = 0; }           // first row touched by the tile's slots
                else { y = r / p.tiles_x; x0 = (r - y * p.tiles_x) * HL_M; }
                for (int c = 0; c < KC; ++c, ++it) {
                    mbar_wait(bar_aempty + 8 * s, ph ^ 1);
                    HL_DBG(0, it);
                    mbar_expect_tx(bar_afull + 8 * s, (uint32_t)n_rows * (p.row64 ? 64 : 128));
                    const int c0 = c * (p.in_split ? 2 * HL_BK : HL_BK);      // element offset of the slice (fp32 or halfs)
                    if (p.s2d) {
                        const int py = c / p.kc_per_py;                         // row phase; the slice is 32 (px, c) values of it
                        hl_tma_load_5d(base + s * p.act_stage, &tmX, bar_afull + 8 * s, (c - py * p.kc_per_py) * HL_BK, py, x0, y, b);
                    } else if (p.row_loads == 1) {
                        tma_load_4d(base + s * p.act_stage, &tmX, bar_afull + 8 * s, c0, x0 - p.pad, y - p.pad, b);
                    } else {
#pragma unroll
                        for (int r2 = 0; r2 < HL_BH; ++r2)
                            tma_load_4d(base + s * p.act_stage + r2 * p.bw * 128, &tmX, bar_afull + 8 * s, c0, x0 - p.pad,
                                        y + (r2 - 1) * p.dil, b);
                    }
                    if (++s == AS) { s = 0; ph ^= 1; }
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
                // one barrier per slice (the last one takes every slice beyond HL_W_STAGES): the MMAs of slice 0 start as
                // soon as ITS nine images have landed, not after the whole set (5k clk at 128 -> 128)
                for (int c = 0; c < KC; ++c) {
                    if (c < HL_W_STAGES) mbar_expect_tx(bar_wfull + 8 * c, (uint32_t)(p.n_taps * (c == HL_W_STAGES - 1 ? KC - c : 1)) * bytes);
                    const uint32_t bar_c = bar_wfull + 8 * (c < HL_W_STAGES ? c : HL_W_STAGES - 1);
                    for (int tap = 0; tap < p.n_taps; ++tap)
                        bulk_load_1d(w_base + (c * p.n_taps + tap) * p.w_stage_bytes, p.w + (size_t)(p.tap_id[tap] * KC + c) * bytes, bytes, bar_c);
                }
            } else {
                for (int j = 0; j < n_items; ++j) {
                    const bool tail = j >= p.rounds;
                    const int ch0 = tail ? tail_ch0 : 0, cn = tail ? p.cn_split : p.Cout;
                    const uint32_t part = (uint32_t)cn * 64;          // rows ch0 .. ch0 + cn of the W_h and of the W_l image
                    for (int c = 0; c < KC; ++c) {
                        for (int tap = 0; tap < p.n_taps; ++tap, ++wt) {
                            const int s = wt & (p.w_stages - 1);
                            mbar_wait(bar_wempty + 8 * s, ((wt >> p.w_shift) & 1) ^ 1);
                            mbar_expect_tx(bar_wfull + 8 * s, 2 * part);
                            const uint8_t* img = p.w + (size_t)(p.tap_id[tap] * KC + c) * bytes + (size_t)ch0 * 64;
                            const uint32_t dst = w_base + s * p.w_stage_bytes;
                            if (cn != p.Cout) {
                                bulk_load_1d(dst, img, part, bar_wfull + 8 * s);
                                bulk_load_1d(dst + part, img + p.b_bytes, part, bar_wfull + 8 * s);
                            } else {
                                bulk_load_1d(dst, img, bytes, bar_wfull + 8 * s);
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            IssueCtx cx;
            cx.idesc_n = (1u << 4) | ((uint32_t)(p.Cout >> 3) << 17) | ((uint32_t)(HL_M >> 4) << 24);
            cx.idesc_w = (1u << 4) | ((uint32_t)((2 * p.Cout) >> 3) << 17) | ((uint32_t)(HL_M >> 4) << 24);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) cx.tapoff[tap] = (uint32_t)(p.tap_dy[tap] * p.bw + p.tap_dx[tap]) * (p.row64 ? 4 : 8);   // 128- or 64-byte rows, >> 4
            cx.wsb16 = (uint32_t)p.w_stage_bytes >> 4;
            cx.cout = (uint32_t)p.Cout;
            cx.bar_wfull = bar_wfull; cx.bar_wempty = bar_wempty;
            cx.w_mask = (uint32_t)p.w_stages - 1; cx.w_shift = (uint32_t)p.w_shift;
            cx.w_lo0 = ((w_base >> 4) & 0x3FFF) | (1u << 16);
            const bool resident = p.w_resident != 0;
            int it = 0, tcount = 0, s = 0;
            uint32_t wt = 0, ph = 0;
            for (int j = 0; j < n_items; ++j, ++tcount) {
                const int t = j < p.rounds ? (int)blockIdx.x + j * (int)gridDim.x : tail_t;
                if (j >= p.rounds && p.cn_split != p.Cout) {          // channel-split tail item: narrower MMAs
                    cx.cout = (uint32_t)p.cn_split;
                    cx.idesc_n = (1u << 4) | ((uint32_t)(p.cn_split >> 3) << 17) | ((uint32_t)(HL_M >> 4) << 24);
                    cx.idesc_w = (1u << 4) | ((uint32_t)((2 * p.cn_split) >> 3) << 17) | ((uint32_t)(HL_M >> 4) << 24);
                }
                const int a = tcount & 1, u = tcount >> 1;
                if (u > 0) {                       // the epilogue has drained this accumulator set
                    mbar_wait(bar_acce + 8 * a, (u - 1) & 1);
                    tc_fence_after();
                }
                const uint32_t d_tile = tmem_acc + a * 256;
                uint32_t c0off = 0;                                   // flat mode: the tile starts c0 slots into its first row
                if (p.flat) { const int r = t % tiles_per_img; c0off = (uint32_t)((r * HL_M) % p.bw) * 8; }
                for (int c = 0; c < KC; ++c, ++it) {
                    if (resident && tcount == 0) mbar_wait(bar_wfull + 8 * (c < HL_W_STAGES ? c : HL_W_STAGES - 1), 0);   // this slice's weights
                    mbar_wait(bar_aconv + 8 * s, ph);
                    HL_DBG(3, it);
                    tc_fence_after();
                    const uint32_t ast = base + s * p.act_stage;
                    const uint32_t a_lo = (((ast >> 4) & 0x3FFF) | (1u << 16)) + c0off;
                    const bool ks2 = p.Cin - c * HL_BK > 16;             // channels 16..31 of the slice are zero padding otherwise
                    const uint32_t first = c == 0 ? 0u : 1u;
                    if (p.row64) {                                       // 16-channel input: one K = 16 step per tap, 64-byte rows
                        if (resident) hl_issue_chunk<true, false, 9, true>(cx, d_tile, a_lo, cx.w_lo0, first, wt);
                        else hl_issue_chunk<false, false, 9, true>(cx, d_tile, a_lo, 0, first, wt);
                    } else if (p.n_taps == 4) {                          // stride 2 as a 2 x 2 convolution (slices are always full)
                        if (resident) hl_issue_chunk<true, true, 4>(cx, d_tile, a_lo, cx.w_lo0 + (uint32_t)c * 4 * cx.wsb16, first, wt);
                        else hl_issue_chunk<false, true, 4>(cx, d_tile, a_lo, 0, first, wt);
                    } else if (resident) {
                        const uint32_t b_lo = cx.w_lo0 + (uint32_t)c * 9 * cx.wsb16;
                        if (ks2) hl_issue_chunk<true, true, 9>(cx, d_tile, a_lo, b_lo, first, wt);
                        else hl_issue_chunk<true, false, 9>(cx, d_tile, a_lo, b_lo, first, wt);
                    } else {
                        if (ks2) hl_issue_chunk<false, true, 9>(cx, d_tile, a_lo, 0, first, wt);
                        else hl_issue_chunk<false, false, 9>(cx, d_tile, a_lo, 0, first, wt);
                    }
                    tc_commit(bar_aempty + 8 * s);
                    HL_DBG(4, it);
                    if (++s == AS) { s = 0; ph ^= 1; }
                }
                tc_commit(bar_accf + 8 * a);
            }
        }
    } else if (warp < 6) {
        // ===================== epilogue (warps 2..5; TMEM lane quadrant = warp % 4) =====================
        const int q = warp & 3;
        const int m = q * 32 + lane;
        uint8_t* stage = base_ptr + p.epi_off + q * 8192;      // this warp's two staging buffers (TMA-store epilogue)
        int nbuf = 0;
        if ((p.tma_y || p.tma_ys) && lane == 0) {
            if (p.tma_y) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY) : "memory");
            if (p.tma_ys) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmYS) : "memory");
            if (p.tma_mask) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmM) : "memory");
        }
        pdl_wait();                                               // mask tiles, accumulate operands and the outputs themselves
        uint8_t* mask_buf = base_ptr + p.epi_off + 4 * 8192 + q * 4096;
        const uint32_t bar_mask = bar_accf + 8 * (4 + q);
        uint32_t mask_phase = 0;
        // one-pass layers on the TMA-store path: software-pipelined epilogue (hl_single_item)
        const bool single = p.single_pass != 0;
        bool pre = false;                                     // the current tile's tcgen05.ld was issued during the previous tile
        uint32_t rm[16], rc[16];
        int tcount = 0;
        // tile coordinates advance by the (constant) tile step without divisions: image b, tile r of the image =
        // (row ty, column tile tx) in row mode
        const int t_step = gridDim.x;
        const int d_b = t_step / tiles_per_img, d_r = t_step - d_b * tiles_per_img;
        const int d_y = d_r / p.tiles_x, d_x = d_r - d_y * p.tiles_x;
        int b = (int)blockIdx.x / tiles_per_img, r = (int)blockIdx.x - b * tiles_per_img;
        int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
        for (int j = 0; j < n_items; ++j, ++tcount) {
            const bool tail = j >= p.rounds;
            const int ch0 = tail ? tail_ch0 : 0, cn = tail ? p.cn_split : p.Cout;
            if (tail) {                                   // the tail item is not on the strided walk
                b = tail_t / tiles_per_img; r = tail_t - b * tiles_per_img;
                ty = r / p.tiles_x; tx = r - ty * p.tiles_x;
            }
            int y, x;
            if (p.flat) { const int slot = r * HL_M + m; y = slot / p.bw; x = slot - y * p.bw; }
            else { y = ty; x = tx * HL_M + m; }
            const int b_now = b;
            r += d_r; b += d_b;
            if (r >= tiles_per_img) { r -= tiles_per_img; ++b; }
            tx += d_x; ty += d_y;
            if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
            if (ty >= p.H) ty -= p.H;                      // row mode: tiles_per_img = tiles_x * H, the image carry is in b
            if (single) {
                // (a channel-split tail item has narrower accumulators: p.single_pass excludes split tails)
                const uint32_t tmem_q = tmem_acc + ((uint32_t)(q * 32) << 16);
                if (p.Cout == 32) hl_single_item32(p, s_bias, tmem_q, tcount, pre, rm, rc, j + 1 < n_items, bar_accf, bar_acce, &tmY, &tmYS, stage, nbuf, lane, x - lane, y, b_now);
                else hl_single_item<16>(p, s_bias, tmem_q, tcount, pre, rm, rc, j + 1 < n_items, bar_accf, bar_acce, &tmY, &tmYS, stage, nbuf, lane, x - lane, y, b_now);
                continue;
            }
            const int a = tcount & 1, u = tcount >> 1;
            if (p.tma_mask && lane == 0) {            // first pass's mask tile: in flight while the MMAs of this tile run
                mbar_expect_tx(bar_mask, (p.Cout & 31) == 0 ? 4096 : 2048);
                tma_load_4d(smem_u32(mask_buf), &tmM, bar_mask, ch0, x - lane, y, b_now);
            }
            mbar_wait(bar_accf + 8 * a, u & 1);
            if (threadIdx.x == 64) HL_DBG(5, tcount);
            tc_fence_after();
            const bool valid = x < p.W && y < p.H;
            const size_t pix = ((size_t)b_now * p.H + y) * p.W + x;
            const bool vec = ((p.y_cs & 3) == 0) && aligned16(p.y) && ((p.cout_valid & 3) == 0) && !p.res &&
                             (!p.mask || (((p.mask_cs & 3) == 0) && aligned16(p.mask)));
            const uint32_t tbase = tmem_acc + ((uint32_t)(q * 32) << 16) + a * 256;
            unsigned long long* edbg = threadIdx.x == 64 ? dbg : nullptr;
            const bool odd = lane & 1;
            const bool valid_o = __shfl_xor_sync(0xffffffffu, (int)valid, 1) != 0;
            const size_t pix_o = (size_t)__shfl_xor_sync(0xffffffffu, (unsigned long long)pix, 1);
            if ((p.Cout & 31) == 0) {
                for (int n0 = 0; n0 < cn; n0 += 32) hl_epilogue_pass<32>(p, s_bias, tbase, n0, valid, pix, valid_o, pix_o, odd, vec, edbg, tcount, &tmY, &tmYS, stage, nbuf, lane, x - lane, y, b_now, n0 + 32 >= cn ? bar_acce + 8 * a : 0u, ch0, cn, &tmM, mask_buf, bar_mask, mask_phase);
            } else {
                for (int n0 = 0; n0 < p.Cout; n0 += 16) hl_epilogue_pass<16>(p, s_bias, tbase, n0, valid, pix, valid_o, pix_o, odd, vec, edbg, tcount, &tmY, &tmYS, stage, nbuf, lane, x - lane, y, b_now, n0 + 16 >= p.Cout ? bar_acce + 8 * a : 0u, 0, p.Cout, &tmM, mask_buf, bar_mask, mask_phase);
            }
            if (threadIdx.x == 64) HL_DBG(6, tcount);
            if (threadIdx.x == 64 && dbg) {
                unsigned long long ns; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
                dbg[80] = clock64(); dbg[81] = tcount + 1; dbg[83] = ns;
                if (tcount == 0) dbg[82] = ns;      // wall clock at the first tile's epilogue
                if (tcount == 0) dbg[84] = dbg[80];
            }
        }
        if ((p.tma_y || p.tma_ys) && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // staging buffers must outlive their reads
    } else {
        // ===================== converters (warps 6..13): fp32 pixel row -> [h | l * 2^11] fp16, in place =====================
        const int ct = threadIdx.x - 192;   // 0..255
        const bool half_rows = p.Cin <= 16 && !p.s2d;
        int it = 0, s = 0;
        uint32_t ph = 0;
        for (int j = 0; j < n_items; ++j) {
            for (int c = 0; c < KC; ++c, ++it) {
                mbar_wait(bar_afull + 8 * s, ph);
                if (ct == 0) HL_DBG(1, it);
                uint8_t* stp = base_ptr + (size_t)s * p.act_stage;
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int R = ct + rr * HL_CONV_THREADS;
                    if (R < n_rows && !p.exp_skip_conv && !p.in_split && p.row64) {
                        // 64-byte rows (16 fp32 channels), 64B swizzle: logical 16-byte chunk j sits at chunk j ^ ((R >> 1) & 3);
                        // in place -> [h: chunks 0-1 | l: chunks 2-3]
                        uint8_t* row = stp + (size_t)R * 64;
                        const int sw = (R >> 1) & 3;
                        float4 v[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) v[j] = *reinterpret_cast<const float4*>(row + ((j ^ sw) << 4));
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            uint4 hq, lq;
                            hl_split8(v[2 * j], v[2 * j + 1], hq, lq);
                            *reinterpret_cast<uint4*>(row + ((j ^ sw) << 4)) = hq;
                            *reinterpret_cast<uint4*>(row + (((j + 2) ^ sw) << 4)) = lq;
                        }
                    } else if (R < n_rows && !p.exp_skip_conv && !p.in_split) {
                        uint8_t* row = stp + (size_t)R * 128;
                        const int sw = R & 7;              // 128B swizzle: logical 16-byte chunk j sits at chunk j ^ (R & 7)
                        if (half_rows) {
                            // 16 input channels (pyramid level 1): channels 16..31 of the slice are the tensor map's zero
                            // fill and the MMAs only take the first K = 16 step -- read 64 bytes, write h to chunks 0-1 and
                            // l to chunks 4-5 (half the converter's shared-memory traffic and arithmetic)
                            float4 v[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) v[j] = *reinterpret_cast<const float4*>(row + ((j ^ sw) << 4));
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                uint4 hq, lq;
                                hl_split8(v[2 * j], v[2 * j + 1], hq, lq);
                                *reinterpret_cast<uint4*>(row + ((j ^ sw) << 4)) = hq;
                                *reinterpret_cast<uint4*>(row + (((j + 4) ^ sw) << 4)) = lq;
                            }
                        } else {
                            float4 v[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] = *reinterpret_cast<const float4*>(row + ((j ^ sw) << 4));
                            uint4 hq[4], lq[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) hl_split8(v[2 * j], v[2 * j + 1], hq[j], lq[j]);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                *reinterpret_cast<uint4*>(row + ((j ^ sw) << 4)) = hq[j];
                                *reinterpret_cast<uint4*>(row + (((j + 4) ^ sw) << 4)) = lq[j];
                            }
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                if (ct == 0) HL_DBG(2, it);
                mbar_arrive(bar_aconv + 8 * s);
                if (++s == AS) { s = 0; ph ^= 1; }
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
                     int in_split, void* y_split, int ys_cs, int s2d) {
    EncodeTiledFn enc = get_encode();
    // Stride 2 (s2d): H, W, Cin arrive as the INPUT's; from here on they describe the half-resolution 2 x 2 problem.
    const int Hin = H, Win = W, Cin_in = Cin;
    if (s2d) {
        if ((H & 1) || (W & 1) || (Cin & 15) || x_cs != Cin || in_split || mask || res || accumulate || dilation != 1) return -1000;
        H /= 2; W /= 2; Cin *= 4;
    }
    if (in_split && (Cin & 31)) return -1000;                   // split rows come in whole 32-channel slices
    if (y_split && ((Cout & 31) || (ys_cs & 15) || !aligned16(y_split))) return -1000;
    if (!enc || Cout > 128 || (Cout & 7) || dilation < 1 || dilation > 16) return -1000;   // two accumulator sets of 2*Cout columns must fit 512
    // rows narrower than a tile: flat mode (d = 1 only) packs several rows into the 128 MMA rows
    const int flat = (W < HL_M && dilation == 1 && !getenv("PWC_HALO_NO_FLAT")) ? 1 : 0;
    const int span = s2d ? 1 : 2 * dilation;                    // extra box columns (and, in rows, box rows - 1) around a tile
    const int fbw = W + span;                                   // flat mode: slots per padded row
    const int nr = flat ? (HL_M - 1 + fbw - 1) / fbw + 1 + span : 0;
    // 16-channel inputs in row-tile mode: 64-byte shared-memory rows (the 128-byte rows would be half zero fill: twice the TMA
    // write traffic and half as many pipeline stages in the same shared memory).  dilation <= 8: the strided 3-row box.
    const int row64 = (Cin == 16 && !flat && !s2d && !in_split && dilation <= 8 && !getenv("PWC_HALO_NO_ROW64")) ? 1 : 0;   // 16->16 at 16x224x512: 118 -> 109 us
    CUtensorMap tmX;
    if (s2d) {
        cuuint64_t dims[5] = {(cuuint64_t)2 * Cin_in, 2, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[4] = {(cuuint64_t)Win * Cin_in * 4, (cuuint64_t)2 * Cin_in * 4, (cuuint64_t)2 * Win * Cin_in * 4,
                                 (cuuint64_t)Hin * Win * Cin_in * 4};
        cuuint32_t box[5] = {HL_BK, 1, (cuuint32_t)(flat ? fbw : HL_M + 1), (cuuint32_t)(flat ? nr : 2), 1};
        cuuint32_t es[5] = {1, 1, 1, 1, 1};
        CUresult r = enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv3x3_tc_halo: cuTensorMapEncodeTiled(x, s2d) failed with %d", (int)r); return PWC_E_BADARG; }
    } else {
        // x_cs counts elements of the input tensor: floats, or halfs of a split tensor (2 * its channel count)
        const cuuint64_t esz = in_split ? 2 : 4;
        cuuint64_t dims[4] = {(cuuint64_t)(in_split ? 2 * Cin : Cin), (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)x_cs * esz, (cuuint64_t)W * x_cs * esz, (cuuint64_t)H * W * x_cs * esz};
        // rows y-d, y, y+d through the traversal stride of the row dimension (TMA allows strides up to 8); larger
        // dilations use one single-row box per row (their row pitch (128+2d)*128 B must keep the 1024-byte swizzle phase)
        const bool strided = dilation <= 8;
        if (!strided && (((HL_M + 2 * dilation) * 128) & 1023)) return -1000;
        cuuint32_t box[4] = {(cuuint32_t)(in_split ? 2 * HL_BK : (row64 ? 16 : HL_BK)), (cuuint32_t)(HL_M + 2 * dilation), (cuuint32_t)(strided ? HL_BH * dilation : 1), 1};
        cuuint32_t es[4] = {1, 1, (cuuint32_t)(strided ? dilation : 1), 1};
        if (flat) { box[1] = (cuuint32_t)fbw; box[2] = (cuuint32_t)nr; es[2] = 1; }
        CUresult r = enc(&tmX, in_split ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)x, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, row64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv3x3_tc_halo: cuTensorMapEncodeTiled(x) failed with %d", (int)r); return PWC_E_BADARG; }
    }
    HaloParams p{};
    p.bias = bias; p.y = y; p.w = (const uint8_t*)w_packed; p.mask = mask; p.res = res; p.res_cs = res_cs;
    p.y_cs = y_cs; p.mask_cs = mask_cs; p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.cout_valid = cout_valid;
    p.tiles_x = (W + HL_M - 1) / HL_M;
    p.flat = flat; p.nr = nr;
    p.tiles_img = flat ? (H * fbw + HL_M - 1) / HL_M : p.tiles_x * H;
    const long long tiles = (long long)p.tiles_img * B;
    if (tiles >= (1ll << 30)) return -1000;
    p.total_tiles = (int)tiles;
    p.kchunks = (Cin + HL_BK - 1) / HL_BK;
    // Work items (HaloParams): whole rounds of tiles, then the tail split along the output channels when that shortens the
    // last, partly filled round (or the only round of a layer with fewer tiles than SMs) by a useful amount.
    const int sms = sm_count();
    const int G = p.total_tiles < sms ? p.total_tiles : sms;
    p.rounds = p.total_tiles >= sms ? p.total_tiles / sms : 0;
    const int tail = p.total_tiles - p.rounds * G;
    p.n_split = 1;
    if (tail > 0 && p.rounds <= 32 && (Cout & 31) == 0 && cout_valid == Cout && !getenv("PWC_HALO_NO_NSPLIT")) {
        for (int k = 4; k >= 2; --k)
            if (Cout % (32 * k) == 0 && tail * k <= sms) { p.n_split = k; break; }
    }
    p.cn_split = Cout / p.n_split;
    p.tail_items = tail * p.n_split;
    const int grid = p.rounds > 0 ? sms : p.tail_items;
    p.b_bytes = Cout * 64;
    p.w_stage_bytes = (2 * p.b_bytes + 1023) / 1024 * 1024;
    p.accumulate = accumulate; p.alpha = alpha; p.mask_alpha = mask_alpha;
    p.in_split = in_split; p.ys = (__half*)y_split; p.ys_cs = ys_cs;
    p.dil = dilation; p.bw = flat ? fbw : HL_M + span; p.row_loads = dilation <= 8 ? 1 : HL_BH;
    p.s2d = s2d; p.pad = s2d ? 0 : dilation; p.box_rows = s2d ? 2 : HL_BH; p.kc_per_py = s2d ? (2 * Cin_in) / HL_BK : 0;
    p.n_taps = s2d ? 4 : 9;
    for (int t = 0; t < p.n_taps; ++t) {
        if (s2d) { p.tap_dy[t] = t >> 1; p.tap_dx[t] = t & 1; p.tap_id[t] = (t >> 1) * 3 + (t & 1); }   // packed as the top-left 2 x 2 of a 3 x 3 kernel
        else { p.tap_dy[t] = t / 3; p.tap_dx[t] = (t % 3) * dilation; p.tap_id[t] = t; }
    }
    p.row64 = row64;
    p.act_stage = ((flat ? nr : p.box_rows) * p.bw * (row64 ? 64 : 128) + 1023) / 1024 * 1024;
    p.desc_mode = 0;
    // TMA-store epilogue: row tiles (not the flat mode: its tiles hold padding slots between rows), whole passes, plain
    // stores (no dgrad mask / residual / accumulate)
    CUtensorMap tmY, tmYS, tmM;
    memset(&tmY, 0, sizeof(tmY)); memset(&tmYS, 0, sizeof(tmYS)); memset(&tmM, 0, sizeof(tmM));
    // (dgrad: the mask tile comes in through TMA as well and `accumulate` becomes a reduce-add store; PWC_HALO_NO_TMA_DGRAD=1
    //  keeps the register epilogue for those)
    const bool dgrad_ok = (!mask || ((mask_cs & 3) == 0 && aligned16(mask))) && !getenv("PWC_HALO_NO_TMA_DGRAD");
    const bool tma_ok = !flat && (!mask || dgrad_ok) && !res && (!accumulate || dgrad_ok) && !(y_split && (mask || accumulate)) &&
                        cout_valid == Cout && (Cout == 16 || (Cout & 31) == 0) && !getenv("PWC_HALO_NO_TMA_STORE");
    if (tma_ok && y && (y_cs & 3) == 0 && aligned16(y)) {
        const cuuint32_t G = Cout == 16 ? 16 : 32;
        cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)y_cs * 4, (cuuint64_t)W * y_cs * 4, (cuuint64_t)H * W * y_cs * 4};
        cuuint32_t box[4] = {G, 32, 1, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&tmY, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)y, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         G == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv3x3_tc_halo: cuTensorMapEncodeTiled(y) failed with %d", (int)r); return PWC_E_BADARG; }
        p.tma_y = 1;
        p.tma_red = accumulate ? 1 : 0;
        if (mask) {
            cuuint64_t mstrides[3] = {(cuuint64_t)mask_cs * 4, (cuuint64_t)W * mask_cs * 4, (cuuint64_t)H * W * mask_cs * 4};
            r = enc(&tmM, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)mask, dims, mstrides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    G == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { set_error("conv3x3_tc_halo: cuTensorMapEncodeTiled(mask) failed with %d", (int)r); return PWC_E_BADARG; }
            p.tma_mask = 1;
        }
    }
    if (tma_ok && y_split && (Cout & 31) == 0) {
        cuuint64_t dims[4] = {(cuuint64_t)2 * Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)ys_cs * 2, (cuuint64_t)W * ys_cs * 2, (cuuint64_t)H * W * ys_cs * 2};
        cuuint32_t box[4] = {64, 32, 1, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&tmYS, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, y_split, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv3x3_tc_halo: cuTensorMapEncodeTiled(y_split) failed with %d", (int)r); return PWC_E_BADARG; }
        p.tma_ys = 1;
    }
    const size_t epi_bytes = (p.tma_y || p.tma_ys) ? 4 * 8192 + (p.tma_mask ? 4 * 4096 : 0) : 0;
    p.single_pass = ((p.tma_y || p.tma_ys) && !p.tma_mask && !p.tma_red && (Cout == 16 || Cout == 32) && p.n_split == 1 && !(p.tma_y && p.tma_ys) && !getenv("PWC_HALO_NO_PIPE_EPI")) ? 1 : 0;
    if (const char* e = getenv("PWC_HALO_EXP")) { p.exp_skip_conv = atoi(e) == 1; p.exp_skip_store = atoi(e) == 2; p.exp_direct_store = atoi(e) == 3; }
    if (const char* e = getenv("PWC_HALO_DESC")) p.desc_mode = atoi(e);
    // (Rotating accumulator sets over the taps were tried in round 1 and again in round 2 with a separate A_l x W_h chain:
    // no gain -- the issue loop, not the accumulator latency, was the limit; profiles/r02_halo_epilogue.log.)
    p.act_stages = 2;   // measured: a third activation stage does not help (174.6 vs 171.5 us at 128->128), the limit is operand bandwidth
    int want_stages = 0;
    if (const char* e = getenv("PWC_HALO_STAGES")) want_stages = atoi(e);
    if (want_stages == 3) p.act_stages = 3;
    const size_t w_all = (size_t)p.n_taps * p.kchunks * p.w_stage_bytes;
    // (a channel-split tail needs [W_h | W_l] sub-images side by side: streamed, not cut out of resident full images)
    p.w_resident = (2 * (size_t)p.act_stage + w_all + epi_bytes + 1024 <= HL_SMEM_BUDGET) && p.n_split == 1 && !getenv("PWC_HALO_NO_RESIDENT");
    if (p.w_resident) {
        // small layers (resident weights): a stage is held from the TMA issue to the last MMA that reads it (~5k clk
        // at 16->16: 1.9k TMA latency + 1.4k conversion + 1.7k MMAs), so two stages cap the tile period at ~2.5k clk
        p.act_stages = 2;
        const int fit = (int)((HL_SMEM_BUDGET - 1024 - w_all - epi_bytes) / p.act_stage);
        const int lim = want_stages >= 2 && want_stages <= HL_MAX_ACT_STAGES ? want_stages : HL_MAX_ACT_STAGES;
        if (fit > 2) p.act_stages = fit < lim ? fit : lim;
    }
    // streaming layers: the weight ring takes what the activation stages leave (4..16 images in flight: with small images --
    // narrow or channel-split layers -- four are not enough to cover the L2 latency of one image per tap)
    p.w_stages = HL_W_MIN_STAGES; p.w_shift = 2;
    if (!p.w_resident) {
        while (p.w_stages < HL_W_STAGES && (size_t)p.act_stages * p.act_stage + (size_t)2 * p.w_stages * p.w_stage_bytes + epi_bytes + 1024 <= HL_SMEM_BUDGET) {
            p.w_stages *= 2; ++p.w_shift;
        }
    }
    size_t smem = (size_t)p.act_stages * p.act_stage + (p.w_resident ? w_all : (size_t)p.w_stages * p.w_stage_bytes) + epi_bytes + 1024;
    if (smem > HL_SMEM_BUDGET) { p.act_stages = 2; smem = (size_t)2 * p.act_stage + (size_t)p.w_stages * p.w_stage_bytes + epi_bytes + 1024; }
    if (smem > HL_SMEM_BUDGET && epi_bytes) { smem -= epi_bytes; p.tma_y = p.tma_ys = p.tma_mask = p.tma_red = p.single_pass = 0; }    // no room for the staging buffers: plain stores
    if (smem > HL_SMEM_BUDGET) return -1000;
    p.epi_off = (int)(smem - 1024 - ((p.tma_y || p.tma_ys) ? epi_bytes : 0));               // 1024-byte aligned: every part before it is
    cudaError_t e = cudaFuncSetAttribute(conv3x3_tc_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("conv3x3_tc_halo: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    static unsigned long long* dbg_buf = nullptr;
    if (getenv("PWC_HALO_DEBUG")) {
        if (!dbg_buf) cudaMalloc(&dbg_buf, 148 * 128 * 8);
        cudaMemsetAsync(dbg_buf, 0, 148 * 128 * 8, st);
        p.dbg = dbg_buf;
    }
    {
        cudaError_t le = launch_pdl(conv3x3_tc_halo_kernel, dim3((unsigned)grid), dim3(HL_THREADS), smem, st, tmX, tmY, tmYS, tmM, p);
        if (le != cudaSuccess) { set_error("conv3x3_tc_halo_kernel: launch: %s", cudaGetErrorString(le)); return (int)le; }
    }
    PWC_CHECK_LAUNCH("conv3x3_tc_halo_kernel");
    if (p.dbg) {   // debugging aid only (synchronises): timeline of the first chunks / tiles of one CTA
        cudaStreamSynchronize(st);
        static int printed = 0;
        if (printed++ < 1) {
            unsigned long long h[128];
            const char* names[10] = {"act_tma", "full_seen", "conv_done", "mma_start", "mma_issued", "acc_seen", "epi_done", "ld_done", "math_done", "stores_out"};
            cudaMemcpy(h, p.dbg + 128 * (grid / 2), 128 * 8, cudaMemcpyDeviceToHost);
            fprintf(stderr, "[halo dbg] cta %d Cin %d Cout %d kchunks %d resident %d (clk from first TMA issue; columns = chunks / tiles 0..7)\n", grid / 2, Cin, Cout, p.kchunks, p.w_resident);
            for (int e = 0; e < 10; ++e) {
                fprintf(stderr, "   %-10s", names[e]);
                for (int t = 0; t < 8; ++t) fprintf(stderr, " %7lld", (long long)(h[e * 8 + t] - h[0]));
                fprintf(stderr, "\n");
            }
            fprintf(stderr, "   last epilogue done at %lld clk after %lld tiles; first -> last epilogue: %lld clk in %lld ns = %.3f GHz\n",
                    (long long)(h[80] - h[0]), (long long)h[81], (long long)(h[80] - h[84]), (long long)(h[83] - h[82]),
                    (double)(h[80] - h[84]) / (double)(h[83] - h[82]));
            // per-CTA totals: slowest CTA
            unsigned long long* all = (unsigned long long*)malloc((size_t)grid * 128 * 8);
            cudaMemcpy(all, p.dbg, (size_t)grid * 128 * 8, cudaMemcpyDeviceToHost);
            long long worst = 0, best = 1ll << 60; int wc = 0;
            for (int c = 0; c < grid; ++c) { const long long d = (long long)(all[c * 128 + 80] - all[c * 128]); if (d > worst) { worst = d; wc = c; } if (d < best) best = d; }
            fprintf(stderr, "   CTA life (first TMA -> last epilogue): min %lld max %lld clk (cta %d, %lld tiles)\n", best, worst, wc, (long long)all[wc * 128 + 81]);
            free(all);
        }
    }
    return 0;
}

}  // namespace pwc
