// 3x3 convolution on tcgen05 with the error-compensated "3 x fp16, scaled residual" scheme (fp32-class).
//
// Same implicit GEMM / pipeline as conv_tc.cu (replaces the same reference node groups,
// modules.py:62-67, 266-274, 306-323), but the operands are split as
//        x = h + l * 2^-11,   h = fp16_rn(x),   l = fp16_rn((x - h) * 2^11)
// (Ootomo & Yokota's residual scaling keeps l out of the fp16 subnormal range), and
//        D = A_h.W_h  +  2^-11 * (A_l.W_h + A_h.W_l)
// is accumulated in two fp32 TMEM accumulator groups with tcgen05.mma.kind::f16 (K = 16 per instruction).
// h.h products are exact in fp32 (11 x 11 bits); the representation error of h + l/2048 is 2^-23 |x|, the
// dropped l.l term 2^-22: fp32-class results, like 3xTF32, at HALF the MMA instructions (6 instead of 12 per
// 32-channel slice) and 2/3 of the shared-memory operand traffic -- on B200 the tf32 variant is bound by
// shared-memory bandwidth (UMMA operand reads + the converter + TMA writes), not by the tensor pipe.
//
// Stage layout: [A raw fp32 16K (TMA, 128B swizzle) | A_h fp16 8K | A_l fp16 8K | W_h fp16 N*64 | W_l fp16 N*64];
// the fp16 tiles are K-major with 64-byte rows (64B swizzle).  The four converter/epilogue warps turn the
// raw fp32 tile into A_h / A_l between TMA arrival and MMA issue; weights are split once on the host side
// of the ABI (pwc_conv3x3_pack_weights_f16).  fp16 range: |x| < 65504 is required (activations and weights
// of this network are O(1e-3 .. 1e2)); values that overflow saturate to inf like any fp16 path would.
#include "tc_common.cuh"
#include <cuda_fp16.h>
#include <cstdlib>

namespace pwc {

constexpr int F16_BM = 128, F16_TW = 16, F16_TH = 8, F16_BK = 32;
constexpr int F16_CONV_THREADS = 256;                 // 8 converter warps (the fp32 -> fp16 hi/lo split is the
                                                      // per-stage bottleneck with 4: ~1000 clk per 16 KB tile)
constexpr int F16_THREADS = 64 + F16_CONV_THREADS;    // + TMA producer warp + MMA warp
constexpr uint32_t F16_A_RAW = F16_BM * F16_BK * 4;   // 16 KB
constexpr uint32_t F16_A_HALF = F16_BM * F16_BK * 2;  // 8 KB
constexpr float F16_SCALE = 2048.f, F16_INV_SCALE = 1.f / 2048.f;

struct F16Params {
    const float* bias; float* y; const uint8_t* w;
    int y_cs, B, H, W, Cin, Cout, dil;
    int OH, OW, stride, pad_t, pad_l;
    int tiles_x, tiles_y, kchunks, total_tiles;
    float alpha;
    int b_bytes;       // channels per CTA * 64 (one fp16 weight tile in shared memory)
    int n_parts, cn, b_bytes_full;   // coarse levels (fewer tiles than SMs): the output channels are split over n_parts CTAs per tile,
                                     // cn channels each (a multiple of 16); b_bytes_full = Cout * 64 = the packed image's W_h -> W_l distance
    int stage_bytes, stages, tmem_cols, n_main, prefetch, shift;
    int k_parts;       // split-K over a cluster of k_parts CTAs (1 or 3: one kernel row of taps each); rank 0 reduces the partial
                       // sums the other ranks leave in its shared memory (DSMEM) and runs the epilogue
    unsigned long long* dbg;   // optional per-CTA timeline (clock64), 8 slots per CTA; nullptr in production
    // dgrad use (pwc_conv3x3_tc_f16_dgrad): bias may be null, stores are limited to the first cout_valid channels
    // (Cout is the MMA N, padded to a multiple of 16), the result is multiplied by leaky'(mask) and optionally
    // accumulated into y
    const float* mask; int mask_cs; float mask_alpha; int accumulate; int cout_valid;
    const float* res; int res_cs;   // residual added after the activation (flow heads, modules.py:275-277, 326)
};

__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// One 32-lane x 16-column block of the result: correction group (scaled by 2^-11) first, then the main accumulators, main 0 last.
__device__ __forceinline__ void f16_load_acc(uint32_t tbase, int n_main, int CN, float (&acc)[16]) {
    uint32_t r[16];
    tmem_ld16(tbase + n_main * CN, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(r[j]) * F16_INV_SCALE;
    for (int a2 = n_main - 1; a2 >= 0; --a2) {
        tmem_ld16(tbase + a2 * CN, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] += __uint_as_float(r[j]);
    }
}

__global__ void __launch_bounds__(F16_THREADS, 2)
conv3x3_tc_f16_kernel(const __grid_constant__ CUtensorMap tmX, const F16Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    __shared__ __align__(8) uint64_t bars[3 * 8 + 1];   // full[8], conv[8], empty[8], acc_full
    __shared__ uint32_t tmem_base_slot;
    __shared__ float s_bias[256 + 16];                  // bias (or zeros): a global load per pass would sit on the epilogue's critical path

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.stages;
    unsigned long long* dbg = p.dbg ? p.dbg + (size_t)blockIdx.x * 80 : nullptr;
    if (dbg && threadIdx.x == 0) dbg[0] = clock64();
    const uint32_t bar_full = smem_u32(&bars[0]), bar_conv = smem_u32(&bars[8]), bar_empty = smem_u32(&bars[16]);
    const uint32_t bar_acc = smem_u32(&bars[24]);

    const int KP = p.k_parts;
    const int krank = KP > 1 ? (int)(blockIdx.x % KP) : 0;   // == %cluster_ctarank (cluster dims (KP,1,1))
    const int bid = KP > 1 ? (int)(blockIdx.x / KP) : (int)blockIdx.x;
    int t = bid / p.n_parts;
    const int ch0 = (bid - t * p.n_parts) * p.cn;     // first output channel of this CTA
    const int CN = p.cn;
    const int tx = t % p.tiles_x; t /= p.tiles_x;
    const int ty = t % p.tiles_y; const int b = t / p.tiles_y;
    const int x0 = tx * F16_TW, y0 = ty * F16_TH;
    // K iterations [IT0, IT0 + KT) of the 9 * kchunks (tap, slice) pairs: all of them, or one kernel row per cluster rank
    const int KT = (KP > 1 ? 3 : 9) * p.kchunks;
    const int IT0 = krank * KT;

    // (p.shift: experiment knob -- start the fp16 A tiles `shift` bytes into their 512-byte swizzle atom to test that
    //  UMMA and the converter agree on an ABSOLUTE-address swizzle; 1 KB of slack follows the A_l tile)
    const uint32_t off_ah = F16_A_RAW + p.shift, off_al = F16_A_RAW + F16_A_HALF + p.shift;
    constexpr uint32_t off_bh = F16_A_RAW + 2 * F16_A_HALF + 1024;
    const uint32_t off_bl = off_bh + p.b_bytes;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_conv + 8 * s, F16_CONV_THREADS);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int c = threadIdx.x; c < 256 + 16; c += F16_THREADS) s_bias[c] = (p.bias && c < p.cout_valid) ? __ldg(p.bias + c) : 0.f;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(p.tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = tmem_base_slot;
    if (dbg && threadIdx.x == 0) dbg[1] = clock64();   // setup done
    pdl_trigger();                                     // (common.cuh) programmatic dependent launch

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
            pdl_wait();                                // activations of the previous layer
            const uint32_t tx_bytes = F16_A_RAW + 2 * p.b_bytes;
            const int PF = p.prefetch;   // stages of L2 prefetch distance for the activation boxes
            for (int it = 0; it < PF && it < KT; ++it) {
                const int tap = (IT0 + it) / p.kchunks, kc = (IT0 + it) - tap * p.kchunks, ky = tap / 3, kx = tap - ky * 3;
                tma_prefetch_4d(&tmX, kc * F16_BK, x0 * p.stride - p.pad_l + kx * p.dil, y0 * p.stride - p.pad_t + ky * p.dil, b);
            }
            // ring stage / phase and the (tap, slice) pairs of the load and of the prefetch as counters: the integer divisions
            // they replace (~40 instructions each) were a visible part of every K iteration of this single thread
            int s = 0, kc = 0, kx = 0, ky = krank;            // IT0 = krank * 3 * kchunks: tap row krank, first tap, first slice
            uint32_t ph = 0;
            int pkc = (IT0 + PF) % p.kchunks, ptap = (IT0 + PF) / p.kchunks, pky = ptap / 3, pkx = ptap - pky * 3;
            for (int it = 0; it < KT; ++it) {
                if (it + PF < KT) {
                    tma_prefetch_4d(&tmX, pkc * F16_BK, x0 * p.stride - p.pad_l + pkx * p.dil, y0 * p.stride - p.pad_t + pky * p.dil, b);
                    if (++pkc == p.kchunks) { pkc = 0; if (++pkx == 3) { pkx = 0; ++pky; } }
                }
                mbar_wait(bar_empty + 8 * s, ph ^ 1);
                const uint32_t st = base + s * p.stage_bytes;
                if (dbg && it >= 8 && it < 24) dbg[8 + (it - 8)] = clock64();
                mbar_expect_tx(bar_full + 8 * s, tx_bytes);
                tma_load_4d(st, &tmX, bar_full + 8 * s, kc * F16_BK, x0 * p.stride - p.pad_l + kx * p.dil,
                            y0 * p.stride - p.pad_t + ky * p.dil, b);
                // [h tile | l tile] of this (tap, slice): one linear bulk copy
                if (p.n_parts > 1) {   // rows ch0 .. ch0 + cn of the W_h and of the W_l image
                    const uint8_t* img = p.w + (size_t)(IT0 + it) * 2 * p.b_bytes_full + (size_t)ch0 * 64;
                    bulk_load_1d(st + off_bh, img, p.b_bytes, bar_full + 8 * s);
                    bulk_load_1d(st + off_bh + p.b_bytes, img + p.b_bytes_full, p.b_bytes, bar_full + 8 * s);
                } else {
                    bulk_load_1d(st + off_bh, p.w + (size_t)(IT0 + it) * 2 * p.b_bytes, 2 * p.b_bytes, bar_full + 8 * s);
                }
                if (++kc == p.kchunks) { kc = 0; if (++kx == 3) { kx = 0; ++ky; } }
                if (++s == S) { s = 0; ph ^= 1; }
            }
            if (dbg) dbg[2] = clock64();   // last TMA issued
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            // kind::f16: D = f32 (bit 4), A = B = F16 (format 0), K-major, N>>3 at [17,23), M>>4 at [24,29)
            const uint32_t idesc = (1u << 4) | ((uint32_t)(CN >> 3) << 17) | ((uint32_t)(F16_BM >> 4) << 24);
            // 64-byte-row K-major tiles: 8-row atoms 512 bytes apart, SWIZZLE_64B (layout type 4)
            const uint64_t desc_hi = ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
            const uint32_t d_corr = tmem_acc + p.n_main * CN;
            int s = 0, am = 0;
            uint32_t ph = 0;
            for (int it = 0; it < KT; ++it) {
                mbar_wait(bar_conv + 8 * s, ph);
                if (dbg && it == 0) dbg[3] = clock64();   // first stage converted
                if (dbg && it >= 8 && it < 24) dbg[56 + (it - 8)] = clock64();
                tc_fence_after();
                const uint32_t st = base + s * p.stage_bytes;
                const uint32_t ah = (((st + off_ah) >> 4) & 0x3FFF) | (1u << 16);
                const uint32_t al = (((st + off_al) >> 4) & 0x3FFF) | (1u << 16);
                const uint32_t bh = (((st + off_bh) >> 4) & 0x3FFF) | (1u << 16);
                const uint32_t bl = (((st + off_bl) >> 4) & 0x3FFF) | (1u << 16);
                const uint32_t d_main = tmem_acc + am * CN;
#pragma unroll
                for (int k = 0; k < F16_BK / 16; ++k)   // K = 16 fp16 = 32 bytes per instruction
                    tc_mma_f16(d_main, desc_hi | (ah + 2 * k), desc_hi | (bh + 2 * k), idesc, (it >= p.n_main || k > 0) ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < F16_BK / 16; ++k)
                    tc_mma_f16(d_corr, desc_hi | (al + 2 * k), desc_hi | (bh + 2 * k), idesc, (it | k) != 0 ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < F16_BK / 16; ++k)
                    tc_mma_f16(d_corr, desc_hi | (ah + 2 * k), desc_hi | (bl + 2 * k), idesc, 1u);
                tc_commit(bar_empty + 8 * s);
                if (++s == S) { s = 0; ph ^= 1; }
                if (++am == p.n_main) am = 0;
            }
            tc_commit(bar_acc);
            if (dbg) dbg[4] = clock64();   // last MMA issued
        }
    } else {
        // ===================== converter, then epilogue =====================
        const int ct = threadIdx.x - 64;   // 0..255
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0; it < KT; ++it) {
            mbar_wait(bar_full + 8 * s, ph);
            if (dbg && ct == 0 && it >= 8 && it < 24) dbg[24 + (it - 8)] = clock64();
            if (dbg && ct == 0 && it == 12) dbg[72] = clock64();
            const uint8_t* stp = base_ptr + (size_t)s * p.stage_bytes;
            const float4* a = reinterpret_cast<const float4*>(stp);
#pragma unroll
            for (int i = 0; i < 1024 / F16_CONV_THREADS; ++i) {
                const int chunk = ct + F16_CONV_THREADS * i;   // physical 16-byte chunk of the raw tile
                const int m = chunk >> 3;                 // pixel row 0..127
                const int lk = (chunk & 7) ^ (m & 7);     // logical 4-channel group (undo the 128B swizzle)
                const float4 v = a[chunk];
                const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                const __half2 l01 = __floats2half2_rn((v.x - f01.x) * F16_SCALE, (v.y - f01.y) * F16_SCALE);
                const __half2 l23 = __floats2half2_rn((v.z - f23.x) * F16_SCALE, (v.w - f23.y) * F16_SCALE);
                // destination in the 64-byte-row, 64B-swizzled fp16 tile: 16-byte chunk j = lk >> 1, half lk & 1
                const uint32_t row = p.shift + m * 64;   // byte offset of the row inside the 1024-aligned tile region
                const uint32_t off = m * 64 + ((((lk >> 1) ^ ((row >> 7) & 3))) << 4) + ((lk & 1) << 3);
                uint2 hv, lv;
                hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
                lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
                *reinterpret_cast<uint2*>(const_cast<uint8_t*>(stp) + off_ah + off) = hv;
                *reinterpret_cast<uint2*>(const_cast<uint8_t*>(stp) + off_al + off) = lv;
            }
            if (dbg && ct == 0 && it == 12) dbg[73] = clock64();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (dbg && ct == 0 && it == 12) dbg[74] = clock64();
            if (dbg && ct == 0 && it >= 8 && it < 24) dbg[40 + (it - 8)] = clock64();
            mbar_arrive(bar_conv + 8 * s);
            if (dbg && ct == 0 && it == 12) dbg[75] = clock64();
            if (++s == S) { s = 0; ph ^= 1; }
        }
        // ---- split-K, ranks 1..KP-1: warps 2..5 leave their partial sums in rank 0's shared memory (DSMEM), laid out
        //      [rank-1][channel group of 4][pixel row] as float4 (consecutive lanes -> consecutive 16 bytes)
        if (warp < 6 && krank != 0) {
            mbar_wait(bar_acc, 0);
            tc_fence_after();
            const int q = warp & 3;
            const int m = q * 32 + lane;
            uint32_t remote;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(base + (uint32_t)S * p.stage_bytes), "r"(0));
            remote += (uint32_t)(krank - 1) * (uint32_t)CN * 512u + (uint32_t)m * 16u;
            for (int n0 = 0; n0 < CN; n0 += 16) {
                float acc[16];
                f16_load_acc(tmem_acc + ((uint32_t)(q * 32) << 16) + n0, p.n_main, CN, acc);
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(remote + (uint32_t)((n0 + j) >> 2) * 2048u),
                                 "f"(acc[j]), "f"(acc[j + 1]), "f"(acc[j + 2]), "f"(acc[j + 3]) : "memory");
            }
        }
    }
    if (KP > 1) {   // every thread of every CTA of the cluster: partial sums are visible to rank 0 afterwards
        __syncwarp();
        asm volatile("barrier.cluster.arrive.release;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
    }
    // ---- epilogue (rank 0): warps 2..5 (TMEM lane quadrant = warp % 4); the other converter warps are done
    if (warp >= 2 && warp < 6 && krank == 0) {
        {
            pdl_wait();                                // mask / residual / accumulate operands
            mbar_wait(bar_acc, 0);
            if (dbg && threadIdx.x == 64) dbg[5] = clock64();   // accumulator complete
            tc_fence_after();
            const int q = warp & 3;
            const int m = q * 32 + lane;
            const int oy = y0 + m / F16_TW, ox = x0 + (m % F16_TW);
            const bool valid = oy < p.OH && ox < p.OW;
            float* yrow = p.y + (((size_t)b * p.OH + oy) * p.OW + ox) * p.y_cs + ch0;
            const float* mrow = p.mask ? p.mask + (((size_t)b * p.OH + oy) * p.OW + ox) * p.mask_cs + ch0 : nullptr;
            const float* rrow = p.res ? p.res + (((size_t)b * p.OH + oy) * p.OW + ox) * p.res_cs + ch0 : nullptr;
            const int cvalid = p.cout_valid - ch0;                 // valid channels of this CTA's range (may exceed CN)
            const bool vec = ((p.y_cs & 3) == 0) && aligned16(p.y) && ((p.cout_valid & 3) == 0) && !p.res &&
                             (!p.mask || (((p.mask_cs & 3) == 0) && aligned16(p.mask)));
            const float4* part = reinterpret_cast<const float4*>(base_ptr + (size_t)S * p.stage_bytes) + m;
            for (int n0 = 0; n0 < CN; n0 += 16) {
                float acc[16];
                f16_load_acc(tmem_acc + ((uint32_t)(q * 32) << 16) + n0, p.n_main, CN, acc);
                for (int r = 1; r < KP; ++r) {          // partial sums of the other kernel rows, in rank order
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 v = part[((size_t)(r - 1) * (CN >> 2) + ((n0 + j) >> 2)) * 128];
                        acc[j] += v.x; acc[j + 1] += v.y; acc[j + 2] += v.z; acc[j + 3] += v.w;
                    }
                }
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[j] += s_bias[ch0 + n0 + j];
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[j] = leaky(acc[j], p.alpha);
                    if (vec) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            if (n0 + j >= cvalid) break;
                            float4 v = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
                            if (mrow) {
                                const float4 m = ldg4(mrow + n0 + j);
                                v.x *= m.x > 0.f ? 1.f : p.mask_alpha; v.y *= m.y > 0.f ? 1.f : p.mask_alpha;
                                v.z *= m.z > 0.f ? 1.f : p.mask_alpha; v.w *= m.w > 0.f ? 1.f : p.mask_alpha;
                            }
                            float4* dst = reinterpret_cast<float4*>(yrow + n0 + j);
                            if (p.accumulate) { const float4 o = *dst; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                            *dst = v;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            if (n0 + j >= cvalid) break;
                            float v = acc[j];
                            if (mrow) v *= __ldg(mrow + n0 + j) > 0.f ? 1.f : p.mask_alpha;
                            if (rrow) v += __ldg(rrow + n0 + j);
                            if (p.accumulate) v += yrow[n0 + j];
                            yrow[n0 + j] = v;
                        }
                    }
                }
            }
        }
    }
    if (dbg && threadIdx.x == 64) dbg[6] = clock64();   // epilogue stores issued
    __syncwarp();   // lanes 1..31 of the producer / MMA warps wait for their lane 0 before the block barrier
    tc_fence_before();
    __syncthreads();
    if (dbg && threadIdx.x == 0) dbg[7] = clock64();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(p.tmem_cols));
    }
}

// HWIO (3,3,Cin,Cout) fp32 -> fp16 weight tiles stored as SHARED-MEMORY IMAGES: for every (tap, 32-channel slice)
// one contiguous block [h tile | l tile], each tile Cout rows x 64 bytes with the 64-byte swizzle already applied
// (16-byte chunk j of row n sits at chunk j ^ ((n >> 1) & 3)).  A pipeline stage then needs ONE linear
// cp.async.bulk of 2*Cout*64 bytes instead of two tensor boxes with 64-byte rows (measured 265 vs 415 clk / 16 KB).
__global__ void pack_weights_f16_kernel(const float* __restrict__ w, __half* __restrict__ out, int Cin, int Cout, int Cin_pad) {
    const size_t total = (size_t)9 * Cout * Cin_pad;
    const int kchunks = Cin_pad / F16_BK;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int c = idx % Cin_pad; size_t r = idx / Cin_pad;
        const int n = r % Cout; const int tap = r / Cout;
        __half h = __float2half_rn(0.f), l = h;
        if (c < Cin) {
            const float v = w[((size_t)tap * Cin + c) * Cout + n];
            h = __float2half_rn(v);
            l = __float2half_rn((v - __half2float(h)) * F16_SCALE);
        }
        const int kc = c / F16_BK, cc = c % F16_BK;
        const size_t tile = ((size_t)(tap * kchunks + kc) * 2) * Cout * F16_BK;          // halfs
        const size_t off = (size_t)n * F16_BK + ((((cc >> 3) ^ ((n >> 1) & 3))) << 3) + (cc & 7);
        out[tile + off] = h;
        out[tile + (size_t)Cout * F16_BK + off] = l;
    }
}

static inline int f16_cin_pad(int Cin) { return (Cin + F16_BK - 1) / F16_BK * F16_BK; }

// Batched form: one launch packs every layer of the model (a training step re-packs ~115 weight tensors: the forward
// kernels after the Adam update, the rotated dgrad kernels before the backward pass; as single launches that was
// 0.7 ms of 16 ms).  Job = 8 x int64 {w, out, K channels, N channels, mode, a, b, c}:
//   mode 0: forward kernel, w = (3,3,K,a) HWIO with a <= N output channels (zero-padded to N: the 2-channel heads);
//   mode 1: stride-1 dgrad kernel of w = (3,3,a,K) HWIO for the input-channel range [b, b+c) zero-padded to N:
//           out(tap, k, n) = w(8 - tap, b + n, k)   (what pwc_conv3x3_rot_weights + pack produce).
__global__ void pack_weights_f16_batched_kernel(const long long* __restrict__ jobs) {
    const long long* j = jobs + 8 * (size_t)blockIdx.y;
    const float* __restrict__ w = reinterpret_cast<const float*>(j[0]);
    __half* __restrict__ out = reinterpret_cast<__half*>(j[1]);
    const int K = (int)j[2], N = (int)j[3], mode = (int)j[4], a = (int)j[5], b = (int)j[6], c = (int)j[7];
    const int Kp = (K + F16_BK - 1) / F16_BK * F16_BK, kchunks = Kp / F16_BK;
    const size_t total = (size_t)9 * N * Kp;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int k = idx % Kp; size_t r = idx / Kp;
        const int n = r % N; const int tap = r / N;
        float v = 0.f;
        if (k < K) {
            if (mode == 0) { if (n < a) v = w[((size_t)tap * K + k) * a + n]; }
            else if (n < c) v = w[((size_t)(8 - tap) * a + b + n) * K + k];
        }
        const __half h = __float2half_rn(v);
        const __half l = __float2half_rn((v - __half2float(h)) * F16_SCALE);
        const int kc = k / F16_BK, cc = k % F16_BK;
        const size_t tile = ((size_t)(tap * kchunks + kc) * 2) * N * F16_BK;          // halfs
        const size_t off = (size_t)n * F16_BK + ((((cc >> 3) ^ ((n >> 1) & 3))) << 3) + (cc & 7);
        out[tile + off] = h;
        out[tile + (size_t)N * F16_BK + off] = l;
    }
}

}  // namespace pwc

extern "C" int pwc_conv3x3_pack_weights_f16_batched(const long long* jobs, int n_jobs, void* stream) {
    using namespace pwc;
    PWC_REQUIRE(jobs, PWC_E_BADARG, "pack_weights_f16_batched: null pointer");
    PWC_REQUIRE(n_jobs > 0 && n_jobs <= 65535, PWC_E_BADARG, "pack_weights_f16_batched: 1..65535 jobs");
    pack_weights_f16_batched_kernel<<<dim3(16, n_jobs), 256, 0, (cudaStream_t)stream>>>(jobs);
    PWC_CHECK_LAUNCH("pack_weights_f16_batched_kernel");
    return 0;
}

extern "C" long long pwc_conv3x3_packed_bytes_f16(int Cin, int Cout) {
    if (Cin <= 0 || Cout <= 0) return 0;
    return 2LL * 9 * Cout * pwc::f16_cin_pad(Cin) * 2;
}

extern "C" int pwc_conv3x3_pack_weights_f16(const float* w_hwio, void* w_packed, int Cin, int Cout, void* stream) {
    using namespace pwc;
    PWC_REQUIRE(w_hwio && w_packed, PWC_E_BADARG, "pack_weights_f16: null pointer");
    PWC_REQUIRE(Cin > 0 && Cout > 0, PWC_E_BADARG, "pack_weights_f16: bad dims");
    pack_weights_f16_kernel<<<148 * 4, 256, 0, (cudaStream_t)stream>>>(w_hwio, (__half*)w_packed, Cin, Cout, f16_cin_pad(Cin));
    PWC_CHECK_LAUNCH("pack_weights_f16_kernel");
    return 0;
}

namespace pwc {
struct F16Extra { const float* mask; int mask_cs; float mask_alpha; int accumulate; int cout_valid; const float* res; int res_cs; };
// conv_tc_halo.cu: halo-resident variant for stride 1, dilation 1, wide rows, Cout <= 128
int launch_conv_halo(const float* x, int x_cs, const void* w_packed, const float* bias, float* y, int y_cs,
                     int B, int H, int W, int Cin, int Cout, int dilation, float alpha, const float* mask, int mask_cs,
                     float mask_alpha, int accumulate, int cout_valid, const float* res, int res_cs, cudaStream_t st,
                     int in_split = 0, void* y_split = nullptr, int ys_cs = 0, int s2d = 0);
}

static int launch_conv_f16(const float* x, int x_cs, const void* w_packed, const float* bias,
                           float* y, int y_cs, int B, int H, int W, int Cin, int Cout, int stride, int dilation,
                           float alpha, const pwc::F16Extra& ex, void* stream) {
    using namespace pwc;
    PWC_REQUIRE(x && w_packed && y, PWC_E_BADARG, "conv3x3_tc_f16: null pointer");
    PWC_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && dilation >= 1, PWC_E_BADARG, "conv3x3_tc_f16: bad dims");
    PWC_REQUIRE(stride == 1 || stride == 2, PWC_E_BADARG, "conv3x3_tc_f16: stride must be 1 or 2");
    PWC_REQUIRE(Cout % 16 == 0 && Cout <= 256, PWC_E_BADARG, "conv3x3_tc_f16: Cout must be a multiple of 16, <= 256");
    PWC_REQUIRE(Cin >= 16, PWC_E_BADARG, "conv3x3_tc_f16: Cin must be >= 16");
    PWC_REQUIRE(x_cs >= Cin && y_cs >= ex.cout_valid && ex.cout_valid > 0 && ex.cout_valid <= Cout, PWC_E_BADARG,
                "conv3x3_tc_f16: channel stride smaller than channel count");
    PWC_REQUIRE(aligned16(x) && (x_cs % 4 == 0) && aligned16(w_packed), PWC_E_ALIGN,
                "conv3x3_tc_f16: x / w_packed must be 16-byte aligned and x_cs a multiple of 4");
    EncodeTiledFn enc = get_encode();
    PWC_REQUIRE(enc != nullptr, PWC_E_NOTBUILT, "conv3x3_tc_f16: cuTensorMapEncodeTiled not available from the driver");
    {
        // halo-resident kernel (conv_tc_halo.cu): one 3 x (128+2d) pixel box per 32-channel slice instead of nine shifted
        // tiles; rows narrower than 128 pixels are packed several to a tile (flat mode).  PWC_CONV_HALO=0 disables it.
        const char* he = getenv("PWC_CONV_HALO");
        const int halo_on = he ? atoi(he) : 1;
        static const int halo_min_w = []() { const char* e = getenv("PWC_HALO_MINW"); return e ? atoi(e) : 8; }();
        if (halo_on && stride == 1 && dilation <= 16 && W >= halo_min_w && Cout <= 128) {
            const int rc = launch_conv_halo(x, x_cs, w_packed, bias, y, y_cs, B, H, W, Cin, Cout, dilation, alpha, ex.mask, ex.mask_cs,
                                            ex.mask_alpha, ex.accumulate, ex.cout_valid, ex.res, ex.res_cs, (cudaStream_t)stream);
            if (rc != -1000) return rc;
        }
    }

    const int cpad = f16_cin_pad(Cin);
    const int OH = (H + stride - 1) / stride, OW = (W + stride - 1) / stride;
    int pad_t = ((OH - 1) * stride + 2 * dilation + 1 - H); pad_t = pad_t > 0 ? pad_t / 2 : 0;
    int pad_l = ((OW - 1) * stride + 2 * dilation + 1 - W); pad_l = pad_l > 0 ? pad_l / 2 : 0;
    const int tiles_x = (OW + F16_TW - 1) / F16_TW, tiles_y = (OH + F16_TH - 1) / F16_TH;
    const long long tiles = (long long)tiles_x * tiles_y * B;
    PWC_REQUIRE(tiles < (1LL << 30), PWC_E_BADARG, "conv3x3_tc_f16: too many tiles");
    CUtensorMap tmX;
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)x_cs * 4, (cuuint64_t)W * x_cs * 4, (cuuint64_t)H * W * x_cs * 4};
        cuuint32_t box[4] = {F16_BK, (cuuint32_t)(F16_TW * stride), (cuuint32_t)(F16_TH * stride), 1};
        cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
        CUresult r = enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)x, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        PWC_REQUIRE(r == CUDA_SUCCESS, PWC_E_BADARG, "conv3x3_tc_f16: cuTensorMapEncodeTiled(x) failed with %d", (int)r);
    }
    F16Params p{};
    p.bias = bias; p.y = y; p.w = (const uint8_t*)w_packed; p.y_cs = y_cs; p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.dil = dilation;
    p.OH = OH; p.OW = OW; p.stride = stride; p.pad_t = pad_t; p.pad_l = pad_l;
    p.tiles_x = tiles_x; p.tiles_y = tiles_y; p.total_tiles = (int)tiles;
    p.kchunks = cpad / F16_BK;
    p.alpha = alpha;
    p.mask = ex.mask; p.mask_cs = ex.mask_cs; p.mask_alpha = ex.mask_alpha; p.accumulate = ex.accumulate; p.cout_valid = ex.cout_valid;
    p.res = ex.res; p.res_cs = ex.res_cs;
    // channel split for the coarse pyramid levels (16 tiles of 128 pixels at level 6): the serial K loop of a CTA streams
    // 2 * Cout * 64 bytes of weights per (tap, slice) and issues N = Cout MMAs while 130 SMs idle
    p.n_parts = 1;
    p.k_parts = 1;
    if (tiles * 2 <= sm_count() && Cout >= 64 && !getenv("PWC_TC_NO_NSPLIT")) {
        for (int k = (int)(sm_count() / tiles); k >= 2; --k)
            if (Cout % k == 0 && (Cout / k) % 16 == 0 && Cout / k >= 32) { p.n_parts = k; break; }
    }
    const int budget = 220 * 1024;
    // everything that follows from the split: channels per CTA, stage size, ring depth, accumulator layout.  Returns the dynamic
    // shared memory of a CTA (0: does not fit).
    auto configure = [&](int n_parts, int k_parts) -> size_t {
        p.n_parts = n_parts; p.k_parts = k_parts;
        p.cn = Cout / p.n_parts;
        p.b_bytes_full = Cout * 64;
        const int Cl = p.cn;
        p.b_bytes = Cl * 64;
        p.stage_bytes = (int)(F16_A_RAW + 2 * F16_A_HALF) + 1024 + 2 * p.b_bytes;
        p.shift = 0;
        if (const char* e = getenv("PWC_TC_SHIFT")) p.shift = atoi(e) & 0x3C0;
        p.stage_bytes = (p.stage_bytes + 1023) / 1024 * 1024;
        // Occupancy beats pipeline depth here (measured, profiles/r01_f16_occupancy.log: 2 CTAs/SM with 2 stages each
        // are 1.38x faster than 1 CTA with 4 stages): a CTA's prologue (descriptor fetch, first TMA latency) and
        // epilogue (TMEM drain + stores, ~15% of its life) overlap the other CTA's main loop.  So: <= 110 KB of
        // shared memory and <= 256 TMEM columns per CTA; short K loops (pyramid layers) get 3 CTAs per SM.
        const int kt = 9 * p.kchunks;
        p.stages = ((kt <= 18 ? 72 : 110) * 1024) / p.stage_bytes;
        if (p.stages < 2) p.stages = 2;
        if (p.stages > 4) p.stages = 4;
        // small problems (coarse pyramid levels: fewer tiles than SMs) cannot use a second CTA per SM anyway and are bound by
        // the serial K loop: give the one CTA a deeper ring so the TMA latency is hidden
        const int part_bytes = p.k_parts > 1 ? (p.k_parts - 1) * p.cn * 512 : 0;   // partial tiles of ranks 1.., next to the ring
        if (tiles <= 148 && !getenv("PWC_TC_NO_DEEP")) { int deep = (200 * 1024 - part_bytes) / p.stage_bytes; if (deep > 6) deep = 6; if (deep > p.stages) p.stages = deep; }
        if (const char* e = getenv("PWC_TC_STAGES")) { int v = atoi(e); if (v >= 2 && v <= 8 && v * p.stage_bytes + part_bytes <= budget) p.stages = v; }
        if (p.stages < 2 || p.stages * p.stage_bytes + part_bytes > budget) return 0;
        p.prefetch = 0;   // L2 prefetch of upcoming activation boxes: measured no gain (profiles/r01_f16_prefetch.log)
        if (const char* e = getenv("PWC_TC_PREFETCH")) p.prefetch = atoi(e);
        p.n_main = 256 / Cl - 1;   // main accumulators the K loop rotates over (+1 correction accumulator)
        if (p.n_main > 3) p.n_main = 3;
        if (p.n_main < 1) p.n_main = 1;
        if (const char* e = getenv("PWC_TC_NMAIN")) { int v = atoi(e); if (v >= 1 && v <= p.n_main) p.n_main = v; }
        int cols = 32;
        while (cols < (p.n_main + 1) * Cl) cols *= 2;
        p.tmem_cols = cols;
        return (size_t)p.stages * p.stage_bytes + part_bytes + 1024;
    };
    // split-K for the same sub-wave layers: a CTA's time is its MMA chain (9 * Cin/16 K-steps x 3 MMAs of ~100+ clk whatever N
    // is), so splitting Cout further shortens nothing; a cluster of 3 CTAs takes one kernel row of taps each and rank 0 sums
    // the partial tiles the others leave in its shared memory (pyramid level 6, 192 -> 192 at 16 x 7 x 16: 6 parts of 32 channels
    // -> 3 parts of 64 channels x 3 kernel rows, 54 -> 18 K iterations per CTA).  Taken only if every cluster is resident at once
    // (clusters are placed inside one GPC: not every SM count is reachable).
    // OPT-IN (PWC_TC_KSPLIT=1): measured +0.7 % on the B = 8 forward (level 6: 31 -> 24 us per layer, the launch's fixed costs
    // remain), but whether a layer is split depends on its tile count, i.e. on the batch size, and the split changes the fp32
    // summation order -- with it a batch of 2 no longer reproduces the single pair bit for bit (tests/test_gpu_fullsize.py).
    const int n_split_only = p.n_parts;
    size_t smem = 0;
    const char* ks = getenv("PWC_TC_KSPLIT");
    if (ks && atoi(ks) == 1 && tiles * 3 <= sm_count() && Cout >= 64 && 9 * (cpad / F16_BK) >= 18) {
        for (int n = (int)(sm_count() / (tiles * 3)); n >= 1 && !smem; --n) {
            if (Cout % n != 0 || (Cout / n) % 16 != 0 || Cout / n < 32 || Cout / n > 128) continue;
            const size_t sm = configure(n, 3);
            if (!sm) continue;
            if (cudaFuncSetAttribute(conv3x3_tc_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) { cudaGetLastError(); continue; }
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(tiles * n * 3)); cfg.blockDim = dim3(F16_THREADS); cfg.dynamicSmemBytes = sm;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 3; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int resident = 0;
            if (cudaOccupancyMaxActiveClusters(&resident, conv3x3_tc_f16_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); continue; }
            if ((long long)resident >= tiles * n) smem = sm;
        }
    }
    if (!smem) smem = configure(n_split_only, 1);
    PWC_REQUIRE(smem != 0, PWC_E_BADARG, "conv3x3_tc_f16: tile does not fit in shared memory");
    PWC_REQUIRE(p.n_main >= 1, PWC_E_BADARG, "conv3x3_tc_f16: Cout too large for the accumulator layout");
    static unsigned long long* dbg_buf = nullptr;
    if (getenv("PWC_TC_DEBUG")) {
        if (!dbg_buf) cudaMalloc(&dbg_buf, 80 * 8 * 65536);
        p.dbg = tiles <= 65536 ? dbg_buf : nullptr;
    }
    cudaError_t e = cudaFuncSetAttribute(conv3x3_tc_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("conv3x3_tc_f16: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    if (p.k_parts > 1)
        launch_pdl_cluster(conv3x3_tc_f16_kernel, dim3((unsigned)(tiles * p.n_parts * p.k_parts)), dim3(F16_THREADS), smem, (cudaStream_t)stream,
                           (unsigned)p.k_parts, tmX, p);
    else
        launch_pdl(conv3x3_tc_f16_kernel, dim3((unsigned)(tiles * p.n_parts)), dim3(F16_THREADS), smem, (cudaStream_t)stream, tmX, p);
    PWC_CHECK_LAUNCH("conv3x3_tc_f16_kernel");
    if (p.dbg) {   // debugging aid only (synchronises!): print the timeline of a few CTAs
        cudaStreamSynchronize((cudaStream_t)stream);
        unsigned long long h[8 * 4];
        const long long ids[4] = {0, 147, tiles / 2, tiles - 1};
        for (int i = 0; i < 4; ++i) cudaMemcpy(h + 8 * i, p.dbg + 80 * ids[i], 64, cudaMemcpyDeviceToHost);
        {
            unsigned long long d[80];
            cudaMemcpy(d, p.dbg + 80 * ids[2], 80 * 8, cudaMemcpyDeviceToHost);
            fprintf(stderr, "[tc_f16 dbg] converter thread 0, it 12: full->stores issued %llu clk, fence %llu clk, arrive %llu clk\n", d[73] - d[72], d[74] - d[73], d[75] - d[74]);
            fprintf(stderr, "[tc_f16 dbg] cta %lld per-stage (it: tma_issue full conv_done mma_start), clk from start\n", ids[2]);
            for (int i = 0; i < 16; ++i)
                fprintf(stderr, "   it %2d: %6llu %6llu %6llu %6llu\n", i + 8, d[8 + i] - d[0], d[24 + i] - d[0], d[40 + i] - d[0], d[56 + i] - d[0]);
        }
        for (int i = 0; i < 4; ++i) {
            unsigned long long* t = h + 8 * i;
            fprintf(stderr, "[tc_f16 dbg] cta %lld KT=%d S=%d: setup %llu | lastTMA %llu | firstConv %llu | lastMMAissue %llu | accDone %llu | epiDone %llu | end %llu (clk from start)\n",
                    ids[i], 9 * p.kchunks, p.stages, t[1] - t[0], t[2] - t[0], t[3] - t[0], t[4] - t[0], t[5] - t[0], t[6] - t[0], t[7] - t[0]);
        }
    }
    return 0;
}

extern "C" int pwc_conv3x3_tc_f16_fwd(const float* x, int x_cs, const void* w_packed, const float* bias,
                                      float* y, int y_cs, int B, int H, int W, int Cin, int Cout, int stride, int dilation,
                                      float alpha, void* stream) {
    PWC_REQUIRE(bias, PWC_E_BADARG, "conv3x3_tc_f16: null pointer");
    const pwc::F16Extra ex{nullptr, 0, 1.f, 0, Cout, nullptr, 0};
    return launch_conv_f16(x, x_cs, w_packed, bias, y, y_cs, B, H, W, Cin, Cout, stride, dilation, alpha, ex, stream);
}

// Same conv with the MMA N padded to Cout_pad (a multiple of 16; the packed kernel and the bias are zero-padded to it),
// only the first Cout channels stored, and an optional residual added after the activation: the 2-channel flow heads
// (modules.py:274-277, 325-326) on the tensor cores.
extern "C" int pwc_conv3x3_tc_f16_head(const float* x, int x_cs, const void* w_packed, const float* bias_pad,
                                       const float* residual, int res_cs, float* y, int y_cs, int B, int H, int W, int Cin,
                                       int Cout, int Cout_pad, int dilation, float alpha, void* stream) {
    PWC_REQUIRE(bias_pad, PWC_E_BADARG, "conv3x3_tc_f16_head: null pointer");
    PWC_REQUIRE(Cout > 0 && Cout <= Cout_pad && (!residual || res_cs >= Cout), PWC_E_BADARG, "conv3x3_tc_f16_head: bad channel counts");
    const pwc::F16Extra ex{nullptr, 0, 1.f, 0, Cout, residual, res_cs};
    return launch_conv_f16(x, x_cs, w_packed, bias_pad, y, y_cs, B, H, W, Cin, Cout_pad, 1, dilation, alpha, ex, stream);
}

// Conv2DBackpropInput of a STRIDE-1 conv on the tensor cores: a SAME conv of dy with the 180-degree-rotated,
// transposed kernel (pwc_conv3x3_rot_weights + pwc_conv3x3_pack_weights_f16).  Cdx_pad is the MMA N (multiple
// of 16, the rotated kernel is zero-padded to it), only the first Cdx channels are stored.
extern "C" int pwc_conv3x3_tc_f16_dgrad(const float* dy, int dy_cs, const void* w_rot_packed, float* dx, int dx_cs,
                                        const float* mask, int mask_cs, float mask_alpha, int accumulate,
                                        int B, int H, int W, int Cdy, int Cdx, int Cdx_pad, int dilation, void* stream) {
    PWC_REQUIRE(!mask || mask_cs >= Cdx, PWC_E_BADARG, "conv3x3_tc_f16_dgrad: mask channel stride smaller than Cdx");
    const pwc::F16Extra ex{mask, mask_cs, mask_alpha, accumulate, Cdx, nullptr, 0};
    return launch_conv_f16(dy, dy_cs, w_rot_packed, nullptr, dx, dx_cs, B, H, W, Cdy, Cdx_pad, 1, dilation, 1.f, ex, stream);
}

// Stride-1 conv of a conv -> conv chain with SPLIT activations (conv_tc_halo.cu): the input and/or the output tensor holds,
// per pixel and 32-channel slice, the 128-byte row [h: 32 x fp16 | l: 32 x fp16 scaled by 2^11] instead of 32 floats (the
// same bytes; what the converter warps of the consumer would otherwise produce from the fp32 tensor on every tile).
//   x_split != 0: x points to a split tensor, x_cs = halfs per pixel (>= 2 * Cin), Cin % 32 == 0.
//   y (fp32, may be NULL) and/or y_split (split, may be NULL; Cout % 32 == 0, ys_cs halfs per pixel, 16-byte aligned).
// Results are bit-identical to pwc_conv3x3_tc_f16_fwd on the fp32 tensors.  Cout <= 128, dilation 1..16.
extern "C" int pwc_conv3x3_tc_f16_split_fwd(const void* x, int x_split, int x_cs, const void* w_packed, const float* bias,
                                            float* y, int y_cs, void* y_split, int ys_cs,
                                            int B, int H, int W, int Cin, int Cout, int dilation, float alpha, void* stream) {
    using namespace pwc;
    PWC_REQUIRE(x && w_packed && bias && (y || y_split), PWC_E_BADARG, "conv3x3_tc_f16_split: null pointer");
    PWC_REQUIRE(B > 0 && H > 0 && W > 0 && Cin >= 16 && Cout > 0 && Cout % 16 == 0 && Cout <= 128 && dilation >= 1 && dilation <= 16,
                PWC_E_BADARG, "conv3x3_tc_f16_split: bad dims (stride 1, Cout <= 128, dilation 1..16)");
    PWC_REQUIRE(aligned16(x) && aligned16(w_packed) && (!y || y_cs >= Cout), PWC_E_ALIGN, "conv3x3_tc_f16_split: alignment / strides");
    PWC_REQUIRE(x_split ? (Cin % 32 == 0 && x_cs >= 2 * Cin && x_cs % 8 == 0) : (x_cs >= Cin && x_cs % 4 == 0), PWC_E_BADARG,
                "conv3x3_tc_f16_split: input channel stride");
    PWC_REQUIRE(!y_split || (Cout % 32 == 0 && ys_cs >= 2 * Cout && ys_cs % 16 == 0 && aligned16(y_split)), PWC_E_BADARG,
                "conv3x3_tc_f16_split: split output needs Cout % 32 == 0 and a 32-byte pixel pitch");
    const int rc = launch_conv_halo(static_cast<const float*>(x), x_cs, w_packed, bias, y, y_cs, B, H, W, Cin, Cout, dilation, alpha,
                                    nullptr, 0, 1.f, 0, Cout, nullptr, 0, (cudaStream_t)stream, x_split ? 1 : 0, y_split, ys_cs);
    PWC_REQUIRE(rc != -1000, PWC_E_BADARG, "conv3x3_tc_f16_split: shape not supported by the halo kernel");
    return rc;
}

// ---------------------------------------------------------------------------------------------------------------------
// Stride-2 3x3 convolution (tf.layers.Conv2D(filters,(3,3),(2,2),'same') + bias + leaky: the first convolution of every
// pyramid level, modules.py:62-63) on the halo kernel, as a stride-1 2 x 2 convolution over the space-to-depth view
//     X'[y][x][(py, px, c)] = X[2y + py][2x + px][c]            (read in place through a 5-D tensor map: no copy),
//     Y[y][x] = sum_{dy,dx in {0,1}} W'[dy][dx] . X'[y + dy][x + dx],   W'[dy][dx][(py,px,c)] = W[2dy+py][2dx+px][c] or 0.
// TF 'SAME' with even H, W and stride 2 pads one row / column AFTER the image only, which is the tensor map's zero fill.
// The streaming kernel this replaces loads nine shifted stride-2 boxes per 32-channel slice (one CTA per 128 outputs);
// here a persistent CTA loads each input element once per tile.  Needs a dense input (x_cs == Cin), even H and W,
// Cin % 16 == 0, Cout % 16 == 0, Cout <= 128.
//   pwc_conv3x3_s2d_reindex : (3,3,Cin,Cout) HWIO -> (3,3,4*Cin,Cout) with W' in taps (0..1, 0..1) and zeros elsewhere; the
//                             result goes through pwc_conv3x3_pack_weights_f16 (Cin' = 4*Cin) like any other kernel.
namespace pwc {
__global__ void s2d_reindex_kernel(const float* __restrict__ w, float* __restrict__ out, int C, int Cout) {
    const size_t total = (size_t)9 * 4 * C * Cout;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int n = idx % Cout; size_t r = idx / Cout;
        const int k = r % (4 * C); const int tap = r / (4 * C);
        const int dy = tap / 3, dx = tap % 3;
        const int py = k / (2 * C), px = (k % (2 * C)) / C, c = k % C;
        const int ky = 2 * dy + py, kx = 2 * dx + px;
        out[idx] = (dy < 2 && dx < 2 && ky < 3 && kx < 3) ? w[((size_t)(ky * 3 + kx) * C + c) * Cout + n] : 0.f;
    }
}
}  // namespace pwc

extern "C" int pwc_conv3x3_s2d_reindex(const float* w_hwio, float* w_s2d, int Cin, int Cout, void* stream) {
    using namespace pwc;
    PWC_REQUIRE(w_hwio && w_s2d && Cin > 0 && Cout > 0, PWC_E_BADARG, "conv3x3_s2d_reindex: bad arguments");
    s2d_reindex_kernel<<<148, 256, 0, (cudaStream_t)stream>>>(w_hwio, w_s2d, Cin, Cout);
    PWC_CHECK_LAUNCH("s2d_reindex_kernel");
    return 0;
}

// x: [B, H, W, Cin] dense fp32; w_packed: pwc_conv3x3_pack_weights_f16 of the re-indexed kernel (Cin' = 4 * Cin);
// y (fp32, may be NULL) and/or y_split (split rows, may be NULL; Cout % 32 == 0): [B, H/2, W/2, ...].
extern "C" int pwc_conv3x3_s2_tc_f16_fwd(const float* x, const void* w_packed, const float* bias, float* y, int y_cs,
                                         void* y_split, int ys_cs, int B, int H, int W, int Cin, int Cout, float alpha, void* stream) {
    using namespace pwc;
    PWC_REQUIRE(x && w_packed && bias && (y || y_split), PWC_E_BADARG, "conv3x3_s2_tc_f16: null pointer");
    PWC_REQUIRE(B > 0 && H > 0 && W > 0 && (H & 1) == 0 && (W & 1) == 0 && Cin >= 16 && Cin % 16 == 0 && Cout > 0 && Cout % 16 == 0 && Cout <= 128,
                PWC_E_BADARG, "conv3x3_s2_tc_f16: bad dims (even H and W, Cin % 16 == 0, Cout % 16 == 0, Cout <= 128)");
    PWC_REQUIRE(aligned16(x) && aligned16(w_packed) && (!y || y_cs >= Cout), PWC_E_ALIGN, "conv3x3_s2_tc_f16: alignment / strides");
    PWC_REQUIRE(!y_split || (Cout % 32 == 0 && ys_cs >= 2 * Cout && ys_cs % 16 == 0 && aligned16(y_split)), PWC_E_BADARG,
                "conv3x3_s2_tc_f16: split output needs Cout % 32 == 0 and a 32-byte pixel pitch");
    const int rc = launch_conv_halo(x, Cin, w_packed, bias, y, y_cs, B, H, W, Cin, Cout, 1, alpha, nullptr, 0, 1.f, 0, Cout, nullptr, 0,
                                    (cudaStream_t)stream, 0, y_split, ys_cs, 1);
    PWC_REQUIRE(rc != -1000, PWC_E_BADARG, "conv3x3_s2_tc_f16: shape not supported by the halo kernel");
    return rc;
}
