// 3x3 convolution (stride 1, any dilation, TF 'SAME' padding) + bias + leaky on the 5th-generation
// tensor cores: implicit GEMM with tcgen05.mma.kind::tf32, TMA-staged operands, TMEM accumulators.
//
// Replaces the Conv2D + BiasAdd + Mul + Maximum node groups of the reference's estimator / context /
// pyramid stacks (modules.py:62-67, 266-274, 306-323) for the layers that are genuine dense
// contractions (Cin >= 32, Cout % 16 == 0).
//
// GEMM view per CTA: D[128 pixels x Cout] = sum over (tap, 32-channel slice) A_tap[128 x 32] * W_tap[32 x Cout]
//   * A_tap tile = one 4-D TMA box {32 ch, 16 px, 8 rows, 1 image} of the NHWC input at the tap's
//     (dy,dx)*dilation offset; out-of-image coordinates are zero-filled by TMA = SAME padding; channels
//     beyond Cin are zero-filled too, so concat buffers of any width work.  128B-swizzled, K-major.
//   * W_tap tile = TMA box {32, Cout} of the packed weights [tap][Cout][Cin_pad] (K-major, 128B swizzle).
//   * one elected thread issues tcgen05.mma (M=128, N=Cout, K=8 per instruction), fp32 accumulators live
//     in TMEM; 4 epilogue warps read them back with tcgen05.ld, add bias, apply leaky-relu and store NHWC
//     with an arbitrary channel stride (concat slots).
//
// Precision: kind::tf32 reads fp32 bit patterns and uses the top 19 bits.  n_split = 1 is plain TF32.
// n_split = 3 is the error-compensated "3xTF32" scheme: x = hi + lo with hi = x & 0xffffe000 and
// lo = rn_tf32(x - hi); D += A_hi*W_hi + A_lo*W_hi + A_hi*W_lo.  W_hi/W_lo are precomputed by
// pwc_conv3x3_pack_weights; A_lo is produced on chip by the 4 epilogue warps between TMA arrival and
// MMA issue (generic-proxy writes + fence.proxy.async), A_hi is the raw tile (hardware truncation).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = A_lo converter during the main loop, then epilogue (TMEM lane quadrant = warp % 4).
#include "tc_common.cuh"
#include <cstdlib>

namespace pwc {

constexpr int TC_BM = 128;        // pixels per CTA
constexpr int TC_TW = 16, TC_TH = 8;
constexpr int TC_THREADS = 192;

struct TcParams {
    const float* bias; float* y;
    int y_cs, B, H, W, Cin, Cout, dil;
    int OH, OW, stride, pad_t, pad_l;   // TF SAME geometry (stride 1 or 2)
    int bk;                             // channels per pipeline stage: 32 (128B swizzle) or 16 (64B swizzle, Cin == 16)
    int a_bytes;                        // 128 * bk * 4
    int tiles_x, tiles_y, kchunks;
    float alpha;
    int b_bytes;      // Cout * bk * 4
    int stage_bytes;  // per-stage smem
    int stages;
    int tmem_cols;
    int n_main;       // number of main accumulators the K loop rotates over
    int prefetch;     // L2 prefetch distance in stages
    int cluster;      // CTAs per cluster sharing the weight tiles by TMA multicast (1, 2 or 4)
    int total_tiles;  // real tiles; CTAs beyond it are padding (run the pipeline, store nothing)
};

// ------------------------------------------------------------------------------------------ kernel
template <int NSPLIT, int BK, int MT>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte aligned operand area (dynamic smem base alignment is not guaranteed beyond 16)
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));

    __shared__ __align__(8) uint64_t bars[3 * 8 + 1];   // full[8], conv[8], empty[8], acc_full
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.stages;
    const uint32_t bar_full = smem_u32(&bars[0]), bar_conv = smem_u32(&bars[8]), bar_empty = smem_u32(&bars[16]);
    const uint32_t bar_acc = smem_u32(&bars[24]);

    // tile coordinates
    int t = blockIdx.x;
    const bool real_tile = t < p.total_tiles;
    const int tx = t % p.tiles_x; t /= p.tiles_x;
    const int ty = t % p.tiles_y; const int b = t / p.tiles_y;   // b == B for padding CTAs: TMA zero-fills
    const int CS = p.cluster;
    const uint32_t crank = CS > 1 ? cluster_ctarank() : 0;
    const uint16_t cmask = (uint16_t)((1u << CS) - 1);
    const int x0 = tx * TC_TW, y0 = ty * TC_TH * MT;   // MT vertically stacked 16x8 pixel tiles share each weight tile
    const int KT = 9 * p.kchunks;

    // per-stage layout: [A raw 16K][A lo 16K (NSPLIT==3)][B hi][B lo (NSPLIT==3)]
    constexpr uint32_t A_BYTES = TC_BM * BK * 4;
    constexpr uint32_t off_alo = MT * A_BYTES;
    constexpr uint32_t off_b = (NSPLIT == 3 ? 2 : 1) * MT * A_BYTES;
    const uint32_t off_blo = off_b + p.b_bytes;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_conv + 8 * s, 128);
            mbar_init(bar_empty + 8 * s, CS);
        }
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(p.tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    if (CS > 1) cluster_sync_all();   // peers' barriers are initialised before any multicast / remote arrive
    tc_fence_after();
    const uint32_t tmem_acc = tmem_base_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
            const uint32_t tx_bytes = MT * A_BYTES + (NSPLIT == 3 ? 2 : 1) * p.b_bytes;
            const int PF = p.prefetch;   // L2 prefetch distance (stages) for the activation boxes
            for (int it = 0; it < PF && it < KT; ++it) {
                const int tap = it / p.kchunks, kc = it - tap * p.kchunks, ky = tap / 3, kx = tap - ky * 3;
#pragma unroll
                for (int t = 0; t < MT; ++t)
                    tma_prefetch_4d(&tmX, kc * BK, x0 * p.stride - p.pad_l + kx * p.dil, (y0 + t * TC_TH) * p.stride - p.pad_t + ky * p.dil, b);
            }
            for (int it = 0; it < KT; ++it) {
                const int s = it % S;
                const uint32_t ph = (it / S) & 1;
                if (it + PF < KT) {
                    const int itp = it + PF;
                    const int tap = itp / p.kchunks, kc = itp - tap * p.kchunks, ky = tap / 3, kx = tap - ky * 3;
#pragma unroll
                    for (int t = 0; t < MT; ++t)
                        tma_prefetch_4d(&tmX, kc * BK, x0 * p.stride - p.pad_l + kx * p.dil, (y0 + t * TC_TH) * p.stride - p.pad_t + ky * p.dil, b);
                }
                mbar_wait(bar_empty + 8 * s, ph ^ 1);
                const int tap = it / p.kchunks, kc = it - tap * p.kchunks;
                const int ky = tap / 3, kx = tap - ky * 3;
                const uint32_t st = base + s * p.stage_bytes;
                mbar_expect_tx(bar_full + 8 * s, tx_bytes);
#pragma unroll
                for (int t = 0; t < MT; ++t)
                    tma_load_4d(st + t * A_BYTES, &tmX, bar_full + 8 * s, kc * BK, x0 * p.stride - p.pad_l + kx * p.dil,
                                (y0 + t * TC_TH) * p.stride - p.pad_t + ky * p.dil, b);
                if (CS == 1) {
                    tma_load_3d(st + off_b, &tmW, bar_full + 8 * s, kc * BK, 0, tap);
                    if (NSPLIT == 3) tma_load_3d(st + off_blo, &tmW, bar_full + 8 * s, kc * BK, 0, 9 + tap);
                } else {
                    // this CTA fetches rows [crank*N/CS, (crank+1)*N/CS) of the weight tile once and multicasts
                    // them into every CTA of the cluster (same smem offset, each CTA's own full barrier)
                    const int nrow = p.Cout / CS;
                    const uint32_t sl = crank * nrow * BK * 4;
                    tma_load_3d_mc(st + off_b + sl, &tmW, bar_full + 8 * s, kc * BK, crank * nrow, tap, cmask);
                    if (NSPLIT == 3) tma_load_3d_mc(st + off_blo + sl, &tmW, bar_full + 8 * s, kc * BK, crank * nrow, 9 + tap, cmask);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            // instruction descriptor: D=f32, A=B=tf32, both K-major, N>>3 at [17,23), M>>4 at [24,29)
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.Cout >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            // upper descriptor word: stride byte offset (next 8-row atom), version 1, swizzle mode
            const uint64_t desc_hi = ((uint64_t)((BK == 32 ? 1024 : 512) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)(BK == 32 ? 2 : 4) << 61);
            for (int it = 0; it < KT; ++it) {
                const int s = it % S;
                const uint32_t ph = (it / S) & 1;
                mbar_wait((NSPLIT == 3 ? bar_conv : bar_full) + 8 * s, ph);
                tc_fence_after();
                const uint32_t st = base + s * p.stage_bytes;
                // The tensor core truncates (round-toward-zero) when it adds into the fp32 accumulator, so
                // the error grows with the length of the accumulation chain.  Main products rotate over
                // n_main TMEM accumulators, the two small correction products go to their own accumulator
                // (|corr| ~ 2^-11 |D|, its truncation error is negligible); the epilogue sums them in fp32.
                const int am = it % p.n_main;
                const int n_acc = p.n_main + (NSPLIT == 3 ? 1 : 0);
                // descriptors differ only in the 14-bit start-address field: build the low words once per
                // stage and step them by 32 bytes (>> 4 = 2) per K = 8 slice -- the single issuing thread
                // must spend well under the MMA's ~64 cycles on address arithmetic
                const uint32_t b_lo = (((st + off_b) >> 4) & 0x3FFF) | (1u << 16);
                const uint32_t blo_lo = (((st + off_blo) >> 4) & 0x3FFF) | (1u << 16);
#pragma unroll
                for (int t = 0; t < MT; ++t) {
                    const uint32_t a_lo = (((st + t * A_BYTES) >> 4) & 0x3FFF) | (1u << 16);
                    const uint32_t alo_lo = (((st + off_alo + t * A_BYTES) >> 4) & 0x3FFF) | (1u << 16);
                    const uint32_t d_main = tmem_acc + (t * n_acc + am) * p.Cout;
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k)
                        tc_mma_tf32(d_main, desc_hi | (a_lo + 2 * k), desc_hi | (b_lo + 2 * k), idesc, (it >= p.n_main || k > 0) ? 1u : 0u);
                    if (NSPLIT == 3) {
                        const uint32_t d_corr = tmem_acc + (t * n_acc + p.n_main) * p.Cout;
#pragma unroll
                        for (int k = 0; k < BK / 8; ++k)
                            tc_mma_tf32(d_corr, desc_hi | (alo_lo + 2 * k), desc_hi | (b_lo + 2 * k), idesc, (it | k) != 0 ? 1u : 0u);
#pragma unroll
                        for (int k = 0; k < BK / 8; ++k)
                            tc_mma_tf32(d_corr, desc_hi | (a_lo + 2 * k), desc_hi | (blo_lo + 2 * k), idesc, 1u);
                    }
                }
                if (CS == 1) tc_commit(bar_empty + 8 * s);     // frees the stage when these MMAs have read it
                else tc_commit_mc(bar_empty + 8 * s, cmask);   // ... in every CTA of the cluster (its producer multicasts into ours)
            }
            tc_commit(bar_acc);                   // accumulator complete
        }
    } else {
        // ===================== converter (3xTF32) then epilogue =====================
        const int ct = threadIdx.x - 64;          // 0..127
        if (NSPLIT == 3) {
            for (int it = 0; it < KT; ++it) {
                const int s = it % S;
                const uint32_t ph = (it / S) & 1;
                mbar_wait(bar_full + 8 * s, ph);
                const float4* a = reinterpret_cast<const float4*>(base_ptr + (size_t)s * p.stage_bytes);
                float4* alo = reinterpret_cast<float4*>(base_ptr + (size_t)s * p.stage_bytes + off_alo);
#pragma unroll
                for (int i = 0; i < (int)(MT * A_BYTES / 16 / 128); ++i) {
                    const float4 v = a[ct + 128 * i];
                    float4 l;
                    l.x = tf32_residual(v.x); l.y = tf32_residual(v.y);
                    l.z = tf32_residual(v.z); l.w = tf32_residual(v.w);
                    alo[ct + 128 * i] = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // make generic writes visible to the MMA (async proxy)
                mbar_arrive(bar_conv + 8 * s);
            }
        }
        // ---- epilogue: TMEM -> registers -> bias + leaky -> global (NHWC, channel stride y_cs)
        mbar_wait(bar_acc, 0);
        tc_fence_after();
        const int q = warp & 3;                    // TMEM lane quadrant this warp may access
        const int m = q * 32 + lane;               // accumulator row = pixel within the 16x8 tile
        const bool vec = ((p.y_cs & 3) == 0) && aligned16(p.y);
        const int n_acc = p.n_main + (NSPLIT == 3 ? 1 : 0);
#pragma unroll 1
        for (int t = 0; t < MT; ++t) {
            const int oy = y0 + t * TC_TH + m / TC_TW, ox = x0 + (m % TC_TW);
            const bool valid = real_tile && oy < p.OH && ox < p.OW;
            float* yrow = p.y + (((size_t)b * p.OH + oy) * p.OW + ox) * p.y_cs;
            for (int n0 = 0; n0 < p.Cout; n0 += 16) {
                // issue the loads of all accumulators (n_acc <= 4), wait once, then sum in fp32:
                // correction (last, smallest) + main n_main-1 .. 1 first, main 0 last
                uint32_t r[4][16];
                const uint32_t tbase = tmem_acc + ((uint32_t)(q * 32) << 16) + t * n_acc * p.Cout + n0;
#pragma unroll
                for (int a = 0; a < 4; ++a)
                    if (a < n_acc) tmem_ld16(tbase + a * p.Cout, r[a]);
                tmem_ld_wait();
                float acc[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#pragma unroll
                for (int a = 3; a >= 0; --a)
                    if (a < n_acc) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) acc[j] += __uint_as_float(r[a][j]);
                    }
                if (valid) {
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = leaky(acc[j] + __ldg(p.bias + n0 + j), p.alpha);
                    if (vec) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            *reinterpret_cast<float4*>(yrow + n0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) yrow[n0 + j] = v[j];
                    }
                }
            }
        }
    }
    __syncwarp();   // lanes 1..31 of the producer / MMA warps wait for their lane 0 before the block barrier
    tc_fence_before();
    __syncthreads();
    if (CS > 1) cluster_sync_all();   // no CTA may exit while peers can still arrive on its barriers
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(p.tmem_cols));
    }
}

// Packs HWIO (3,3,Cin,Cout) weights into [2][9][Cout][Cin_pad]: plane 0 = w & 0xffffe000 (tf32-exact),
// plane 1 = rn_tf32(w - plane0), zero for channels >= Cin.
__global__ void pack_weights_kernel(const float* __restrict__ w, float* __restrict__ out, int Cin, int Cout, int Cin_pad) {
    const size_t total = (size_t)9 * Cout * Cin_pad;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int c = idx % Cin_pad; size_t r = idx / Cin_pad;
        const int n = r % Cout; const int tap = r / Cout;
        float hi = 0.f, lo = 0.f;
        if (c < Cin) {
            const float v = w[((size_t)tap * Cin + c) * Cout + n];
            hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
            lo = tf32_residual(v);
        }
        out[idx] = hi;
        out[total + idx] = lo;
    }
}

static inline int tc_bk(int Cin) { return Cin >= 32 ? 32 : 16; }
static inline int cin_pad(int Cin) { const int bk = tc_bk(Cin); return (Cin + bk - 1) / bk * bk; }

}  // namespace pwc

extern "C" long long pwc_conv3x3_packed_bytes(int Cin, int Cout) {
    if (Cin <= 0 || Cout <= 0) return 0;
    return 2LL * 9 * Cout * pwc::cin_pad(Cin) * 4;
}

extern "C" int pwc_conv3x3_pack_weights(const float* w_hwio, float* w_packed, int Cin, int Cout, void* stream) {
    using namespace pwc;
    PWC_REQUIRE(w_hwio && w_packed, PWC_E_BADARG, "pack_weights: null pointer");
    PWC_REQUIRE(Cin > 0 && Cout > 0, PWC_E_BADARG, "pack_weights: bad dims");
    pack_weights_kernel<<<148 * 4, 256, 0, (cudaStream_t)stream>>>(w_hwio, w_packed, Cin, Cout, cin_pad(Cin));
    PWC_CHECK_LAUNCH("pack_weights_kernel");
    return 0;
}

extern "C" int pwc_conv3x3_tc_fwd(const float* x, int x_cs, const float* w_packed, const float* bias,
                                  float* y, int y_cs, int B, int H, int W, int Cin, int Cout, int stride, int dilation,
                                  float alpha, int n_split, void* stream) {
    using namespace pwc;
    PWC_REQUIRE(stride == 1 || stride == 2, PWC_E_BADARG, "conv3x3_tc: stride must be 1 or 2");
    PWC_REQUIRE(x && w_packed && bias && y, PWC_E_BADARG, "conv3x3_tc: null pointer");
    PWC_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && dilation >= 1, PWC_E_BADARG, "conv3x3_tc: bad dims");
    PWC_REQUIRE(n_split == 1 || n_split == 3, PWC_E_BADARG, "conv3x3_tc: n_split must be 1 or 3");
    PWC_REQUIRE(Cout % 16 == 0 && Cout <= 256, PWC_E_BADARG, "conv3x3_tc: Cout must be a multiple of 16, <= 256");
    PWC_REQUIRE(Cin == 16 || Cin >= 32, PWC_E_BADARG, "conv3x3_tc: Cin must be 16 or >= 32");
    PWC_REQUIRE(x_cs >= Cin && y_cs >= Cout, PWC_E_BADARG, "conv3x3_tc: channel stride smaller than channel count");
    PWC_REQUIRE(aligned16(x) && (x_cs % 4 == 0) && aligned16(w_packed), PWC_E_ALIGN,
                "conv3x3_tc: x / w_packed must be 16-byte aligned and x_cs a multiple of 4");
    EncodeTiledFn enc = get_encode();
    PWC_REQUIRE(enc != nullptr, PWC_E_NOTBUILT, "conv3x3_tc: cuTensorMapEncodeTiled not available from the driver");

    const int cpad = cin_pad(Cin);
    // TF 'SAME': out = ceil(in / s); pad_before = max((out-1)*s + 2*d + 1 - in, 0) / 2
    const int OH = (H + stride - 1) / stride, OW = (W + stride - 1) / stride;
    int pad_t = ((OH - 1) * stride + 2 * dilation + 1 - H); pad_t = pad_t > 0 ? pad_t / 2 : 0;
    int pad_l = ((OW - 1) * stride + 2 * dilation + 1 - W); pad_l = pad_l > 0 ? pad_l / 2 : 0;
    // M tiles per CTA: two vertically stacked 16x8 tiles reuse every weight tile twice (the kernel is bound by
    // bytes delivered into the SM, not by L2 or the tensor pipe) when TMEM has room for both accumulator sets.
    // Measured on B200 (profiles/r01_tc_probe_mt.log): +17% in tf32 mode, -5% in 3xtf32 mode (shared-memory
    // bandwidth, not delivered bytes, is the limiter there), so it is opt-in via PWC_TC_MT=2.
    int mt = 1;
    if (const char* e = getenv("PWC_TC_MT")) mt = (atoi(e) == 2 && Cout <= 128 && OH > TC_TH) ? 2 : 1;
    // K slice per stage: 16 channels when two 128-wide tiles would leave too few pipeline stages
    int bk = tc_bk(Cin);
    if (mt == 2 && n_split == 3 && Cout > 64) bk = 16;
    const int tiles_x = (OW + TC_TW - 1) / TC_TW, tiles_y = (OH + TC_TH * mt - 1) / (TC_TH * mt);
    const CUtensorMapSwizzle swz = bk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    const long long tiles = (long long)tiles_x * tiles_y * B;
    PWC_REQUIRE(tiles < (1LL << 30), PWC_E_BADARG, "conv3x3_tc: too many tiles");
    // CTAs per cluster that share each weight tile through TMA multicast (cuts the L2->SM weight traffic).
    // PWC_TC_CLUSTER=2|4 enables it.
    int cluster = 1;   // measured on B200: multicast (2: -4%, 4: -13%) does not pay, the kernel is not L2-bound
    if (const char* e = getenv("PWC_TC_CLUSTER")) cluster = atoi(e);
    if (cluster != 1 && cluster != 2 && cluster != 4) cluster = 1;
    while (cluster > 1 && (Cout / cluster) % 8 != 0) cluster /= 2;
    CUtensorMap tmX, tmW;
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)x_cs * 4, (cuuint64_t)W * x_cs * 4, (cuuint64_t)H * W * x_cs * 4};
        // with element strides s the box spans TW*s x TH*s input pixels and delivers TW x TH of them
        cuuint32_t box[4] = {(cuuint32_t)bk, (cuuint32_t)(TC_TW * stride), (cuuint32_t)(TC_TH * stride), 1};
        cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
        CUresult r = enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)x, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        PWC_REQUIRE(r == CUDA_SUCCESS, PWC_E_BADARG, "conv3x3_tc: cuTensorMapEncodeTiled(x) failed with %d", (int)r);
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)cpad, (cuuint64_t)Cout, 18};
        cuuint64_t strides[2] = {(cuuint64_t)cpad * 4, (cuuint64_t)Cout * cpad * 4};
        cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)(Cout / cluster), 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)w_packed, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        PWC_REQUIRE(r == CUDA_SUCCESS, PWC_E_BADARG, "conv3x3_tc: cuTensorMapEncodeTiled(w) failed with %d", (int)r);
    }
    TcParams p{};
    p.bias = bias; p.y = y; p.y_cs = y_cs; p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.dil = dilation;
    p.tiles_x = tiles_x; p.tiles_y = tiles_y;
    p.cluster = cluster; p.total_tiles = (int)tiles;
    p.prefetch = 0;   // measured no gain
    if (const char* e = getenv("PWC_TC_PREFETCH")) p.prefetch = atoi(e);
    p.OH = OH; p.OW = OW; p.stride = stride; p.pad_t = pad_t; p.pad_l = pad_l;
    p.bk = bk; p.a_bytes = TC_BM * bk * 4;
    p.kchunks = cpad / bk;
    p.alpha = alpha;
    p.b_bytes = Cout * bk * 4;
    p.stage_bytes = (n_split == 3 ? 2 : 1) * (mt * p.a_bytes + p.b_bytes);
    p.stage_bytes = (p.stage_bytes + 1023) / 1024 * 1024;
    const int budget = 220 * 1024;
    p.stages = budget / p.stage_bytes;
    if (p.stages > 8) p.stages = 8;
    const int kt = 9 * p.kchunks;
    // short K loops are dominated by per-tile prologue/epilogue latency: keep smem small so that 3-4 CTAs share
    // an SM and overlap each other's fill and drain phases
    if (kt <= 18) { int cap = (56 * 1024) / p.stage_bytes; if (cap < 2) cap = 2; if (p.stages > cap) p.stages = cap; }
    if (const char* e = getenv("PWC_TC_STAGES")) { int v = atoi(e); if (v >= 2 && v <= 8 && v * p.stage_bytes <= budget) p.stages = v; }
    PWC_REQUIRE(p.stages >= 2, PWC_E_BADARG, "conv3x3_tc: tile does not fit in shared memory");
    // TMEM columns: n_main main accumulators (+1 correction accumulator for 3xTF32), N columns each
    const int extra = n_split == 3 ? 1 : 0;
    p.n_main = 512 / (mt * Cout) - extra;
    if (p.n_main > 3) p.n_main = 3;
    if (n_split == 1) p.n_main = 1;
    PWC_REQUIRE(p.n_main >= 1, PWC_E_BADARG, "conv3x3_tc: Cout too large for the accumulator layout");
    int cols = 32;
    while (cols < mt * (p.n_main + extra) * Cout) cols *= 2;
    p.tmem_cols = cols;
    const size_t smem = (size_t)p.stages * p.stage_bytes + 1024;
    cudaStream_t st = (cudaStream_t)stream;
    void (*kern)(const CUtensorMap, const CUtensorMap, const TcParams);
    if (n_split == 3) {
        if (mt == 2) kern = bk == 32 ? conv3x3_tc_kernel<3, 32, 2> : conv3x3_tc_kernel<3, 16, 2>;
        else kern = bk == 32 ? conv3x3_tc_kernel<3, 32, 1> : conv3x3_tc_kernel<3, 16, 1>;
    } else {
        if (mt == 2) kern = bk == 32 ? conv3x3_tc_kernel<1, 32, 2> : conv3x3_tc_kernel<1, 16, 2>;
        else kern = bk == 32 ? conv3x3_tc_kernel<1, 32, 1> : conv3x3_tc_kernel<1, 16, 1>;
    }
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("conv3x3_tc: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    if (cluster == 1) {
        // plain launch: cudaLaunchKernelEx with a cluster attribute fails under ncu while a CUDA graph is being captured
        kern<<<(unsigned)tiles, TC_THREADS, smem, st>>>(tmX, tmW, p);
        e = cudaGetLastError();
    } else {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)((tiles + cluster - 1) / cluster * cluster), 1, 1);
        cfg.blockDim = dim3(TC_THREADS, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, kern, tmX, tmW, p);
    }
    if (e != cudaSuccess) { set_error("conv3x3_tc: launch: %s", cudaGetErrorString(e)); return (int)e; }
    PWC_CHECK_LAUNCH("conv3x3_tc_kernel");
    return 0;
}
