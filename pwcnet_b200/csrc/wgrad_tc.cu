// Conv2DBackpropFilter on the tcgen05 tensor cores (3 x fp16 split, fp32-class) -- training step, train.py:89.
//
//   dw[tap, ci, co] += sum_{b, y, x} x[b, y*s - pad + ky*d, x*s - pad + kx*d, ci] * dy[b, y, x, co]
//
// Per tap this is a GEMM with M = Cin, N = Cout and K = B*OH*OW output positions.  tcgen05 wants both operands with K
// contiguous, i.e. channel-major planes, so each operand is first transposed and split once:
//   pwc_tsplit_f16   NHWC fp32 (any channel stride) -> fp16 planes (h, l * 2^11) of shape (B, C, H, Wp) -- a
//                    32-channel x 64-position shared-memory transpose; for dy it also reduces the bias gradient; for the
//                    conv input it writes three copies, one per horizontal tap shift, sampled at the output columns
//                    (a 1-position shift is 2 bytes in these planes and TMA box starts must be 16-byte aligned).
//   pwc_conv3x3_wgrad_tc   CTA = (pair of taps, 128-row ci tile, co tile, range of output rows).  Per 32-position
//                    chunk of an output row the producer warp TMA-loads the dy tiles [N rows][32 positions] once and,
//                    for each of its taps, the x tiles [128 ci][32 positions] of copy kx at input row y*s - pad + ky*d
//                    (zero padding above / below through out-of-bounds fill); one thread issues
//                    A_h x [B_h | B_l] (N = 2*Ntile: main | correction) and A_l x B_h per K = 16 step into the tap's
//                    TMEM accumulators, which live for the whole row range; four epilogue warps then add
//                    main + 2^-11 * correction to dw with fp32 atomics (un-permuting concat layers through cin_map).
// K is split over CTAs (row ranges), which also keeps the fp32 accumulation chains short.
#include "tc_common.cuh"
#include <cuda_fp16.h>

namespace pwc {

constexpr float WT_SCALE = 2048.f, WT_INV_SCALE = 1.f / 2048.f;

// ------------------------------------------------------------------------------------------------ transpose + split
// out holds n_shift copies, copy kx = [h plane | l plane], each (B, C, H, OWp):
//   copy_kx[b, c, y, ox] = x[b, y, ox*stride - pad_l + kx*dil, c]   (0 outside the image).
// n_shift = 1, stride = 1, pad_l = 0 is the plain transpose (used for dy, whose per-channel sums go to db); the conv
// input uses n_shift = 3: the horizontal tap shift (2 bytes in these planes) cannot be expressed as a TMA coordinate
// (box starts must be 16-byte aligned), the vertical one can.
__global__ void __launch_bounds__(256) tsplit_kernel(const float* __restrict__ x, int x_cs, __half* __restrict__ out,
                                                     int B, int H, int W, int C, int OW, int OWp, int n_shift, int stride,
                                                     int dil, int pad_l, float* __restrict__ db) {
    __shared__ float tile[32][65];
    const int tid = threadIdx.x;
    const int cb = blockIdx.y * 32;
    const int n_xt = (OW + 63) / 64;
    const long long total = (long long)B * H * n_xt;
    const size_t plane = (size_t)B * C * H * OWp;
    const int wc = tid >> 3, wg = tid & 7;       // write role: channel wc, positions 8*wg .. 8*wg+7
    float bacc = 0.f;
    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
        const int xt = (int)(t % n_xt); const long long r = t / n_xt;
        const int y = (int)(r % H), b = (int)(r / H);
        const int x0 = xt * 64;
        for (int kx = 0; kx < n_shift; ++kx) {
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int e = tid + 256 * j;
                const int px = e >> 3, c4 = e & 7;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                const int c = cb + 4 * c4;
                const int sx = (x0 + px) * stride - pad_l + kx * dil;
                if (x0 + px < OW && sx >= 0 && sx < W && c < C) v = ldg4(x + (((size_t)b * H + y) * W + sx) * x_cs + c);   // C % 4 == 0
                tile[4 * c4 + 0][px] = v.x; tile[4 * c4 + 1][px] = v.y; tile[4 * c4 + 2][px] = v.z; tile[4 * c4 + 3][px] = v.w;
            }
            __syncthreads();
            float s = 0.f;
            if (cb + wc < C && x0 + 8 * wg < OWp) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) { v[i] = tile[wc][8 * wg + i]; s += v[i]; }
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
                    const float2 f = __half22float2(h);
                    const __half2 l = __floats2half2_rn((v[2 * i] - f.x) * WT_SCALE, (v[2 * i + 1] - f.y) * WT_SCALE);
                    hw[i] = *reinterpret_cast<const uint32_t*>(&h); lw[i] = *reinterpret_cast<const uint32_t*>(&l);
                }
                const size_t off = (size_t)(2 * kx) * plane + (((size_t)b * C + cb + wc) * H + y) * OWp + x0 + 8 * wg;
                *reinterpret_cast<uint4*>(out + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                *reinterpret_cast<uint4*>(out + off + plane) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
            if (db) {
                s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
                bacc += s;
            }
        }
    }
    if (db && wg == 0 && cb + wc < C) atomicAdd(db + cb + wc, bacc);
}

// ------------------------------------------------------------------------------------------------ GEMM
constexpr int WT_STAGES = 4;
constexpr int WT_THREADS = 192;            // TMA warp, MMA warp, 4 epilogue warps
constexpr uint32_t WT_A_TILE = 128 * 64;   // 8 KB: 128 ci rows x 32 positions of fp16

struct WtcParams {
    float* dw; const int* cin_map;
    int dw_cin, Cin, Cout, n_tile, co_tiles, ci_tiles;
    int B, OH, OW, stride, dil, pad_t, pad_l;
    int total_rows, rows_per_cta, xchunks;
    int b_bytes;       // n_tile * 64: one fp16 dy tile
    int a_rows;        // rows of an x tile that are actually loaded (min(128, Cin rounded up to 8)); the rest of the
                       // 128-row MMA operand is stale shared memory and only feeds accumulator rows that are never read
    int stage_bytes;
};

__device__ __forceinline__ void wt_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

struct WtcMaps { CUtensorMap a[6]; CUtensorMap b[2]; };   // a[2*kx + plane]: shifted copies of x^T; b[plane]: dy^T

__global__ void __launch_bounds__(WT_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ WtcMaps maps, const WtcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t bars[2 * WT_STAGES + 1];   // full[4], empty[4], acc_full
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[WT_STAGES]), bar_acc = smem_u32(&bars[2 * WT_STAGES]);

    // blockIdx.x = tap group (fastest-varying: the five CTAs that read the same dy rows and x copies are scheduled
    // together, so the re-reads hit L2), blockIdx.z = row range
    const int tg = blockIdx.x;                               // taps 2 tg, 2 tg + 1 (group 4: tap 8 only)
    const int ntaps = tg == 4 ? 1 : 2;
    const int mt = blockIdx.y % p.ci_tiles, nt = blockIdx.y / p.ci_tiles;
    const int ci0 = mt * 128, co0 = nt * p.n_tile;
    const int row_begin = blockIdx.z * p.rows_per_cta;
    const int row_end = min(row_begin + p.rows_per_cta, p.total_rows);
    const int n_chunks = (row_end - row_begin) * p.xchunks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < WT_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_acc, 1);
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
    const uint32_t off_b = 4 * WT_A_TILE;                   // [A_h0 | A_l0 | A_h1 | A_l1 | B_h | B_l]

    if (warp == 0) {
        if (n_chunks > 0 && elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b[0]) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b[1]) : "memory");
            const uint32_t tx_bytes = ntaps * 2 * p.a_rows * 64 + 2 * p.b_bytes;
            int it = 0;
            for (int r = row_begin; r < row_end; ++r) {
                const int b = r / p.OH, y = r - b * p.OH;
                for (int xc = 0; xc < p.xchunks; ++xc, ++it) {
                    const int s = it % WT_STAGES;
                    mbar_wait(bar_empty + 8 * s, ((it / WT_STAGES) & 1) ^ 1);
                    const uint32_t st = base + s * p.stage_bytes;
                    mbar_expect_tx(bar_full + 8 * s, tx_bytes);
                    tma_load_4d(st + off_b, &maps.b[0], bar_full + 8 * s, 32 * xc, y, co0, b);
                    tma_load_4d(st + off_b + p.b_bytes, &maps.b[1], bar_full + 8 * s, 32 * xc, y, co0, b);
                    for (int j = 0; j < ntaps; ++j) {
                        const int tap = 2 * tg + j, ky = tap / 3, kx = tap - ky * 3;
                        const int iy = y * p.stride - p.pad_t + ky * p.dil;      // the horizontal shift is baked into copy kx
                        tma_load_4d(st + (2 * j) * WT_A_TILE, &maps.a[2 * kx], bar_full + 8 * s, 32 * xc, iy, ci0, b);
                        tma_load_4d(st + (2 * j + 1) * WT_A_TILE, &maps.a[2 * kx + 1], bar_full + 8 * s, 32 * xc, iy, ci0, b);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (n_chunks > 0 && elect_one()) {
            const uint32_t idesc_n = (1u << 4) | ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t idesc_w = (1u << 4) | ((uint32_t)((2 * p.n_tile) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            // K-major rows of 64 bytes (32 positions of fp16), 64B swizzle, 8-row groups 512 bytes apart
            const uint64_t desc_hi = ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
            for (int it = 0; it < n_chunks; ++it) {
                const int s = it % WT_STAGES;
                mbar_wait(bar_full + 8 * s, (it / WT_STAGES) & 1);
                tc_fence_after();
                const uint32_t st = base + s * p.stage_bytes;
                const uint64_t bh = desc_hi | (uint64_t)((((st + off_b) >> 4) & 0x3FFF) | (1u << 16));
                for (int j = 0; j < ntaps; ++j) {
                    const uint64_t ah = desc_hi | (uint64_t)((((st + (2 * j) * WT_A_TILE) >> 4) & 0x3FFF) | (1u << 16));
                    const uint64_t al = desc_hi | (uint64_t)((((st + (2 * j + 1) * WT_A_TILE) >> 4) & 0x3FFF) | (1u << 16));
                    const uint32_t d_main = tmem_acc + j * 2 * p.n_tile, d_corr = d_main + p.n_tile;
#pragma unroll
                    for (int k = 0; k < 2; ++k) wt_mma(d_main, ah + 2 * k, bh + 2 * k, idesc_w, (it | k) != 0 ? 1u : 0u);   // A_h x [B_h | B_l]
#pragma unroll
                    for (int k = 0; k < 2; ++k) wt_mma(d_corr, al + 2 * k, bh + 2 * k, idesc_n, 1u);                          // A_l x B_h
                }
                tc_commit(bar_empty + 8 * s);
            }
            tc_commit(bar_acc);
        }
    } else if (n_chunks > 0) {
        // ===================== epilogue (warps 2..5; TMEM lane quadrant = warp % 4): lane = input channel =====================
        const int q = warp & 3;
        const int ci = ci0 + q * 32 + lane;
        mbar_wait(bar_acc, 0);
        tc_fence_after();
        int dst = -1;
        if (ci < p.Cin) dst = p.cin_map ? __ldg(p.cin_map + ci) : ci;
        const uint32_t tbase = tmem_acc + ((uint32_t)(q * 32) << 16);
        for (int j = 0; j < ntaps; ++j) {
            const int tap = 2 * tg + j;
            float* drow = p.dw + ((size_t)tap * p.dw_cin + (dst < 0 ? 0 : dst)) * p.Cout + co0;
            for (int n0 = 0; n0 < p.n_tile; n0 += 16) {
                uint32_t rm[16], rc[16];
                tmem_ld16(tbase + j * 2 * p.n_tile + n0, rm);
                tmem_ld16(tbase + j * 2 * p.n_tile + p.n_tile + n0, rc);
                tmem_ld_wait();
                if (dst >= 0) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (co0 + n0 + i < p.Cout) atomicAdd(drow + n0 + i, __uint_as_float(rm[i]) + __uint_as_float(rc[i]) * WT_INV_SCALE);
                }
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

static inline int wt_wp(int W) { return (W + 7) / 8 * 8; }

// plane (B, C, H, Wp) fp16 with logical width Wl: box {32 positions, 1 row, bc channels, 1}
static bool wt_map(CUtensorMap* tm, const __half* basep, int B, int C, int H, int Wl, int Wp, int bc) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)Wl, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)Wp * 2, (cuuint64_t)H * Wp * 2, (cuuint64_t)C * H * Wp * 2};
    const cuuint32_t box[4] = {32, 1, (cuuint32_t)bc, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(basep), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace pwc

using namespace pwc;

extern "C" long long pwc_tsplit_bytes(int B, int H, int OW, int C, int n_shift) {
    if (B <= 0 || H <= 0 || OW <= 0 || C <= 0 || n_shift <= 0) return 0;
    return 2LL * n_shift * B * C * H * wt_wp(OW) * 2;
}

extern "C" int pwc_tsplit_f16(const float* x, int x_cs, void* out, int B, int H, int W, int C, int n_shift, int stride,
                              int dilation, float* db, void* stream) {
    PWC_REQUIRE(x && out && B > 0 && H > 0 && W > 0 && C > 0, PWC_E_BADARG, "tsplit_f16: bad arguments");
    PWC_REQUIRE((n_shift == 1 || n_shift == 3) && stride >= 1 && dilation >= 1, PWC_E_BADARG, "tsplit_f16: n_shift must be 1 or 3");
    PWC_REQUIRE(n_shift == 3 || (stride == 1 && !0), PWC_E_BADARG, "tsplit_f16: the plain transpose has stride 1");
    PWC_REQUIRE((C & 3) == 0 && (x_cs & 3) == 0 && aligned16(x) && aligned16(out), PWC_E_ALIGN,
                "tsplit_f16: C and x_cs must be multiples of 4, x/out 16-byte aligned");
    // n_shift = 3: positions are the conv's output columns, copy kx samples input column ox*stride - pad_l + kx*dilation
    int OW = W, pad_l = 0;
    if (n_shift == 3) {
        OW = (W + stride - 1) / stride;
        const int tot = (OW - 1) * stride + 2 * dilation + 1 - W;
        pad_l = tot > 0 ? tot / 2 : 0;
    }
    const int OWp = wt_wp(OW);
    const long long tiles = (long long)B * H * ((OW + 63) / 64);
    const int cblocks = (C + 31) / 32;
    PWC_REQUIRE(cblocks <= 65535, PWC_E_BADARG, "tsplit_f16: too many channels");
    long long gx = (148LL * 8 + cblocks - 1) / cblocks;
    if (gx > tiles) gx = tiles;
    if (gx < 1) gx = 1;
    tsplit_kernel<<<dim3((unsigned)gx, cblocks), 256, 0, (cudaStream_t)stream>>>(x, x_cs, (__half*)out, B, H, W, C, OW, OWp, n_shift,
                                                                               stride, dilation, pad_l, n_shift == 1 ? db : nullptr);
    PWC_CHECK_LAUNCH("tsplit_kernel");
    return 0;
}

extern "C" int pwc_conv3x3_wgrad_tc(const void* xT, const void* dyT, float* dw, const int* cin_map, int dw_cin,
                                    int B, int H, int W, int Cin, int Cout, int stride, int dilation, void* stream) {
    PWC_REQUIRE(xT && dyT && dw, PWC_E_BADARG, "conv3x3_wgrad_tc: null pointer");
    PWC_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && dw_cin > 0 && dilation >= 1, PWC_E_BADARG, "conv3x3_wgrad_tc: bad dims");
    PWC_REQUIRE(stride == 1 || stride == 2, PWC_E_BADARG, "conv3x3_wgrad_tc: stride must be 1 or 2");
    PWC_REQUIRE((Cout % 16) == 0, PWC_E_BADARG, "conv3x3_wgrad_tc: Cout must be a multiple of 16");
    PWC_REQUIRE(cin_map || dw_cin == Cin, PWC_E_BADARG, "conv3x3_wgrad_tc: dw_cin must equal Cin without a cin_map");
    WtcParams p{};
    p.dw = dw; p.cin_map = cin_map; p.dw_cin = dw_cin; p.Cin = Cin; p.Cout = Cout;
    p.B = B; p.stride = stride; p.dil = dilation;
    p.OH = (H + stride - 1) / stride; p.OW = (W + stride - 1) / stride;
    int pt = (p.OH - 1) * stride + 2 * dilation + 1 - H; p.pad_t = pt > 0 ? pt / 2 : 0;
    int pl = (p.OW - 1) * stride + 2 * dilation + 1 - W; p.pad_l = pl > 0 ? pl / 2 : 0;
    p.co_tiles = (Cout + 127) / 128;
    p.n_tile = ((Cout + p.co_tiles - 1) / p.co_tiles + 15) / 16 * 16;
    p.ci_tiles = (Cin + 127) / 128;
    p.total_rows = B * p.OH;
    p.xchunks = (p.OW + 31) / 32;
    const int tiles = p.ci_tiles * p.co_tiles;
    int R = (148 * 2 + 5 * tiles - 1) / (5 * tiles);
    if (R > p.total_rows) R = p.total_rows;
    if (R < 1) R = 1;
    p.rows_per_cta = (p.total_rows + R - 1) / R;
    R = (p.total_rows + p.rows_per_cta - 1) / p.rows_per_cta;
    p.b_bytes = p.n_tile * 64;
    p.stage_bytes = (int)(4 * WT_A_TILE) + 2 * p.b_bytes;
    p.stage_bytes = (p.stage_bytes + 1023) / 1024 * 1024;
    WtcMaps maps;
    {
        const int OWp = wt_wp(p.OW);
        const size_t xplane = (size_t)B * Cin * H * OWp, yplane = (size_t)B * Cout * p.OH * OWp;
        const __half* xb = (const __half*)xT;
        const __half* yb = (const __half*)dyT;
        bool ok = true;
        p.a_rows = p.ci_tiles > 1 ? 128 : (Cin + 7) / 8 * 8;
        for (int i = 0; i < 6; ++i) ok = ok && wt_map(&maps.a[i], xb + (size_t)i * xplane, B, Cin, H, p.OW, OWp, p.a_rows);
        for (int i = 0; i < 2; ++i) ok = ok && wt_map(&maps.b[i], yb + (size_t)i * yplane, B, Cout, p.OH, p.OW, OWp, p.n_tile);
        PWC_REQUIRE(ok, PWC_E_BADARG, "conv3x3_wgrad_tc: cuTensorMapEncodeTiled failed");
    }
    const size_t smem = (size_t)WT_STAGES * p.stage_bytes + 1024;
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("conv3x3_wgrad_tc: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    PWC_REQUIRE(R <= 65535, PWC_E_BADARG, "conv3x3_wgrad_tc: too many row ranges");
    dim3 grid(5, tiles, R);
    wgrad_tc_kernel<<<grid, WT_THREADS, smem, (cudaStream_t)stream>>>(maps, p);
    PWC_CHECK_LAUNCH("wgrad_tc_kernel");
    return 0;
}
