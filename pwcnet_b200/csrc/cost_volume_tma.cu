// Cost volume, search range 4, TMA-pipelined persistent kernel (sm_100a) -- the level-2 roofline kernel.
//
//   out[b,y,x,(v+4)*9+(h+4)] = leaky( (1/C) * sum_c f0[b,y,x,c] * f1[b,y+v,x+h,c] ),  zero outside the image
//   (CostVolumeLayer.__call__ / get_cost, reference modules.py:164-204).
//
// Why this shape.  The op is nominally HBM-bound (580 B per pixel at C = 32) but carries 2592 FMAs per pixel:
// at 70 % of the measured HBM peak the FP32 pipe must run at 57 % of ITS peak at the same time, so the kernel is
// organised like an SGEMM: compute warps issue (almost) only FFMA + LDS.128, everything else is off their path.
//
//   * CTA = 4 warps (one per SM sub-partition, up to 255 registers each), 2 CTAs per SM, persistent over
//     7 x 32 pixel tiles.  Lane 0 of warp 0 doubles as the producer (a fifth warp would cap registers at 168).
//   * The producer streams 16-channel slices with TMA (cp.async.bulk.tensor, 4-D boxes straight out of the
//     NHWC tensors): f0 box {16 ch, 32 px, 7 rows}, f1 halo box {16 ch, 41 px, 15 rows} at (x0-4, y0-4); TMA's
//     out-of-bounds zero fill IS the reference's zero padding.  Two stages, full/empty mbarriers: slice k+1
//     (or the next tile's first slice) lands while slice k is consumed; no staging instructions at all.
//   * Work item = (output row y, vertical shift v, 16-pixel strip): 16 x 9 = 144 accumulators per thread, the
//     9-wide f1 window slides along the strip in registers.  7 rows x 9 shifts x 2 strips = 126 items = 4 warps
//     with 98 % of the lanes busy and no wasted (y, v) pairs.  Per 4 channels a thread issues 16 + 24 LDS.128
//     for 576 FFMA (3.6 FMA per shared-memory word).
//   * Bank conflicts are avoided without padding the channel dim (TMA writes dense boxes): a quarter-warp holds
//     ONE output row with v = 0..7 (the last quarter of the odd warp holds v = 8 of the seven rows), lane l visits
//     the four 16-byte channel chunks of a slice in the rotated order ((l & 7) >> 1) + k) & 3, and the f1 box is
//     41 pixels wide so that the pixel parity of a lane's f1 row alternates with v.  The 8 lanes of a quarter then
//     cover all 8 bank groups on every f1 load (measured: 4.0 wavefronts per LDS.128); f0 loads are broadcasts.
//     (Each lane sums the channels in its own fixed order: deterministic, independent of batch neighbours.)
//   * Epilogue: scale + leaky in registers, then a per-warp shared-memory transpose per pixel column; the 324-byte
//     per-pixel runs leave as float4 stores when the destination is 16-byte aligned (concat slots are).
//   * Optional f0 -> concat-slot copy (f0_copy) is a TMA store of the staged f0 box issued by the producer lane.
#include "cost_volume.cuh"
#include "tc_common.cuh"

namespace pwc {

constexpr int T_TY = 7, T_TX = 32, T_SX = 16, T_CH = 16;
constexpr int T_HY = T_TY + 8;                        // 15 f1 rows
constexpr int T_HXB = 41;                             // f1 box width: 40 needed, 41 makes the pixel pitch odd
constexpr int T_F0_BYTES = T_TY * T_TX * T_CH * 4;    // 14336
constexpr int T_F1_BYTES = T_HY * T_HXB * T_CH * 4;   // 39360
constexpr int T_STAGE_BYTES = ((T_F0_BYTES + T_F1_BYTES + 127) / 128) * 128;   // 53760
constexpr int T_NSTAGE = 2;
constexpr int T_CWARPS = 4;
constexpr int T_THREADS = T_CWARPS * 32;              // 128
constexpr int T_XROW = 104;                           // transpose slab row pitch (81 used; = 8 mod 32 keeps the STS conflict-free)
constexpr int T_XPOSE_WORDS = 4 * T_XROW;             // per-warp transpose slab: 4 rows (>= 32 x 9 words)
constexpr int T_BAR_OFF = T_NSTAGE * T_STAGE_BYTES + T_CWARPS * T_XPOSE_WORDS * 4;
constexpr int T_SMEM_BYTES = T_BAR_OFF + 64;
static_assert(T_F0_BYTES % 128 == 0, "f1 box must start 128-byte aligned");
static_assert(2 * (T_SMEM_BYTES + 1024) <= 233472, "two CTAs per SM");

#define FMA_V(acc, a, b) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc) : "f"(a), "f"(b))

struct CvTmaParams {
    float* out;
    int out_cs, B, H, W, n_slices;
    int tiles_x, tiles_y, total_tiles;
    float alpha, inv_c;
    int copy_f0;
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

template <bool VEC>
__global__ void __launch_bounds__(T_THREADS, 2)
cost_volume_tma_kernel(const __grid_constant__ CUtensorMap tm_f0, const __grid_constant__ CUtensorMap tm_f1,
                       const __grid_constant__ CUtensorMap tm_copy, const CvTmaParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const uint32_t smem_base = smem_u32(smem_raw);
    const uint32_t bar_full = smem_base + T_BAR_OFF;          // [2]
    const uint32_t bar_empty = smem_base + T_BAR_OFF + 16;    // [2]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < T_NSTAGE; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, T_CWARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // ================================ consumers ================================
    const int strip = warp >> 1, half = warp & 1;
    // Lane -> (output row y, vertical shift index v).  Every quarter-warp (8 lanes) holds ONE row with v = 0..7,
    // except the last quarter of the odd warp, which holds v = 8 of rows 0..6 (lane 31 is a spare).
    const int q = lane >> 3, l8 = lane & 7;
    const bool tail_q = half == 1 && q == 3;
    const bool item_ok = !(tail_q && l8 == 7);
    const int y = tail_q ? (l8 < 7 ? l8 : 6) : half * 4 + q;
    const int v = tail_q ? 8 : l8;
    const int r = y + v;
    const int rho = l8 >> 1;
    const int a_off = (y * T_TX + strip * T_SX) * T_CH;        // floats into the f0 box
    const int b_off = (r * T_HXB + strip * T_SX) * T_CH;       // floats into the f1 box (col 0 = x0 - 4)
    float* xpose = reinterpret_cast<float*>(smem_raw + T_NSTAGE * T_STAGE_BYTES) + warp * T_XPOSE_WORDS;

    // Producer duties ride on lane 0 of warp 0: at the top of job j it refills the other stage with job j + 1
    // (the stage was last read by job j - 1), so a slice is always in flight while one is consumed.
    const bool is_producer = tid == 0;
    int pt = blockIdx.x, pc = 0;   // (tile, slice) of the next job to load
    auto produce = [&](int pj) {
        if (pt >= p.total_tiles) return;
        const int ps = pj & 1, use = pj >> 1;
        const uint32_t stage = smem_base + ps * T_STAGE_BYTES;
        if (use > 0) mbar_wait(bar_empty + 8 * ps, (use - 1) & 1);
        if (p.copy_f0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        const int ptx = pt % p.tiles_x, pty = (pt / p.tiles_x) % p.tiles_y, pb = pt / (p.tiles_x * p.tiles_y);
        mbar_expect_tx(bar_full + 8 * ps, T_F0_BYTES + T_F1_BYTES);
        tma_load_4d(stage, &tm_f0, bar_full + 8 * ps, pc * T_CH, ptx * T_TX, pty * T_TY, pb);
        tma_load_4d(stage + T_F0_BYTES, &tm_f1, bar_full + 8 * ps, pc * T_CH, ptx * T_TX - 4, pty * T_TY - 4, pb);
        if (++pc == p.n_slices) { pc = 0; pt += gridDim.x; }
    };
    if (is_producer) produce(0);

    int job = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const int tx = t % p.tiles_x, ty = (t / p.tiles_x) % p.tiles_y, b = t / (p.tiles_x * p.tiles_y);
        const int x0 = tx * T_TX, y0 = ty * T_TY;
        const int xs0 = x0 + strip * T_SX;
        const bool active = xs0 < p.W;

        float acc[T_SX][9];
#pragma unroll
        for (int i = 0; i < T_SX; ++i)
#pragma unroll
            for (int j = 0; j < 9; ++j) acc[i][j] = 0.f;

        for (int c = 0; c < p.n_slices; ++c, ++job) {
            const int s = job & 1, ph = (job >> 1) & 1;
            if (is_producer) produce(job + 1);
            __syncwarp();
            mbar_wait(bar_full + 8 * s, ph);
            if (is_producer && p.copy_f0) {
                tma_store_4d(&tm_copy, smem_base + s * T_STAGE_BYTES, c * T_CH, x0, y0, b);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (active) {
                const float* f0s = reinterpret_cast<const float*>(smem_raw + s * T_STAGE_BYTES);
                const float* f1s = reinterpret_cast<const float*>(smem_raw + s * T_STAGE_BYTES + T_F0_BYTES);
#pragma unroll 1
                for (int kk = 0; kk < 4; ++kk) {
                    const int kap = ((kk + rho) & 3) * 4;
                    const float* ap = f0s + a_off + kap;
                    const float* bp = f1s + b_off + kap;
                    float4 w[9];   // sliding window over the f1 row: columns i .. i+8
#pragma unroll
                    for (int q = 0; q < 9; ++q) w[q] = *reinterpret_cast<const float4*>(bp + q * T_CH);
#pragma unroll
                    for (int i = 0; i < T_SX; ++i) {
                        const float4 a = *reinterpret_cast<const float4*>(ap + i * T_CH);
                        // volatile asm pins the PTX order: runs of 9 FFMAs sharing one f0 operand (operand-reuse
                        // cache), 9 independent accumulators between two updates of the same one.
#pragma unroll
                        for (int j = 0; j < 9; ++j) FMA_V(acc[i][j], a.x, w[(i + j) % 9].x);
#pragma unroll
                        for (int j = 0; j < 9; ++j) FMA_V(acc[i][j], a.y, w[(i + j) % 9].y);
#pragma unroll
                        for (int j = 0; j < 9; ++j) FMA_V(acc[i][j], a.z, w[(i + j) % 9].z);
#pragma unroll
                        for (int j = 0; j < 9; ++j) FMA_V(acc[i][j], a.w, w[(i + j) % 9].w);
                        if (i + 1 < T_SX)   // column i leaves the window, column i + 9 enters
                            w[i % 9] = *reinterpret_cast<const float4*>(bp + (i + 9) * T_CH);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + 8 * s);
        }

        if (!active) continue;
        // ---- epilogue: per pixel column i the warp's values go through a shared-memory slab laid out like the
        // global per-pixel runs (slab row = image row, word = channel), then leave as float4 (VEC) or scalar units.
        // Warp 0 of a strip owns channels [0,72) of rows 0..3; warp 1 owns [0,72) of rows 4..6 (slab rows 0..2) and
        // channels [72,81) of all seven rows (rows 0..3 of those sit in slab row 3, 12 words apart).
        float* outb = p.out + (((size_t)b * p.H + y0) * p.W + xs0) * p.out_cs;
        const int cs = p.out_cs;
        const size_t rowpitch = (size_t)p.W * cs;
        constexpr int NU = VEC ? 3 : 9;          // unit slots per lane (float4 or scalar)
        constexpr int UW = VEC ? 4 : 1;          // words per unit
        constexpr int UPR = 72 / UW;             // units per row in the [0,72) part
        constexpr int UPT = VEC ? 2 : 9;         // units per row in the [72,81) tail (+1 scalar when VEC)
        int su[NU]; float* qu[NU];               // slab word / destination in pixel column 0 (nullptr = none)
#pragma unroll
        for (int m = 0; m < NU; ++m) {
            const int u = lane + 32 * m;
            const int nmain = (half == 0 ? 4 : 3) * UPR;
            int row, w, sw;
            bool ok;
            if (u < nmain) { const int rr = u / UPR; w = (u - rr * UPR) * UW; row = half * 4 + rr; sw = T_XROW * rr + w; ok = true; }
            else {
                const int t = u - nmain, rr = t / UPT;
                row = rr; w = 72 + (t - rr * UPT) * UW; ok = half == 1 && rr < 7;
                sw = rr >= 4 ? T_XROW * (rr - 4) + w : 3 * T_XROW + 12 * rr + (w - 72);
            }
            ok = ok && y0 + row < p.H;
            su[m] = ok ? sw : 0;
            qu[m] = ok ? outb + row * rowpitch + w : nullptr;
        }
        int s1 = 0; float* q1 = nullptr;         // VEC only: channel 80 of rows 0..6 (odd warp, lanes 0..6)
        if (VEC && half == 1 && lane < 7 && y0 + lane < p.H) {
            s1 = lane >= 4 ? T_XROW * (lane - 4) + 80 : 3 * T_XROW + 12 * lane + 8;
            q1 = outb + lane * rowpitch + 80;
        }
        float* xs = xpose + (tail_q ? (y >= 4 ? T_XROW * (y - 4) + 72 : 3 * T_XROW + 12 * y) : T_XROW * q + 9 * v);
        const int ncol = min(T_SX, p.W - xs0);
#pragma unroll
        for (int i = 0; i < T_SX; ++i) {
            if (i >= ncol) break;
            if (item_ok) {
#pragma unroll
                for (int j = 0; j < 9; ++j) xs[j] = leaky(acc[i][j] * p.inv_c, p.alpha);
            }
            __syncwarp();
#pragma unroll
            for (int m = 0; m < NU; ++m) {
                if (qu[m]) {
                    if (VEC) *reinterpret_cast<float4*>(qu[m] + (ptrdiff_t)i * cs) = *reinterpret_cast<const float4*>(xpose + su[m]);
                    else qu[m][(ptrdiff_t)i * cs] = xpose[su[m]];
                }
            }
            if (VEC && q1) q1[(ptrdiff_t)i * cs] = xpose[s1];
            __syncwarp();
        }
    }
    if (is_producer && p.copy_f0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

static bool make_map(CUtensorMap* tm, const float* base, int cs, int B, int H, int W, int C, int bx, int by) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)cs * 4, (cuuint64_t)W * cs * 4, (cuuint64_t)H * W * cs * 4};
    const cuuint32_t box[4] = {(cuuint32_t)T_CH, (cuuint32_t)bx, (cuuint32_t)by, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int launch_cv_tma(const CvParams& q, cudaStream_t st) {
    if (q.flow || (q.C % T_CH) != 0 || (q.f0_cs & 3) || (q.f1_cs & 3) || !aligned16(q.f0) || !aligned16(q.f1))
        return CV_TMA_UNSUPPORTED;
    if (q.f0_copy && ((q.f0_copy_cs & 3) || !aligned16(q.f0_copy))) return CV_TMA_UNSUPPORTED;
    if ((long long)q.H * q.W * q.out_cs >= (1ll << 31)) return CV_TMA_UNSUPPORTED;   // 32-bit offsets in the epilogue
    CUtensorMap tm0, tm1, tmc;
    if (!make_map(&tm0, q.f0, q.f0_cs, q.B, q.H, q.W, q.C, T_TX, T_TY) ||
        !make_map(&tm1, q.f1, q.f1_cs, q.B, q.H, q.W, q.C, T_HXB, T_HY))
        return CV_TMA_UNSUPPORTED;
    if (q.f0_copy) {
        if (!make_map(&tmc, q.f0_copy, q.f0_copy_cs, q.B, q.H, q.W, q.C, T_TX, T_TY)) return CV_TMA_UNSUPPORTED;
    } else {
        tmc = tm0;
    }
    CvTmaParams p{};
    p.out = q.out; p.out_cs = q.out_cs; p.B = q.B; p.H = q.H; p.W = q.W; p.n_slices = q.C / T_CH;
    p.tiles_x = (q.W + T_TX - 1) / T_TX; p.tiles_y = (q.H + T_TY - 1) / T_TY;
    p.total_tiles = p.tiles_x * p.tiles_y * q.B;
    p.alpha = q.alpha; p.inv_c = q.inv_c; p.copy_f0 = q.f0_copy ? 1 : 0;
    const bool vec = aligned16(q.out) && (q.out_cs & 3) == 0;
    auto kern = vec ? cost_volume_tma_kernel<true> : cost_volume_tma_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T_SMEM_BYTES);
    if (e != cudaSuccess) { set_error("cost_volume_tma: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    const int cap = 2 * sm_count();
    const int grid = p.total_tiles < cap ? p.total_tiles : cap;
    kern<<<grid, T_THREADS, T_SMEM_BYTES, st>>>(tm0, tm1, tmc, p);
    PWC_CHECK_LAUNCH("cost_volume_tma_kernel");
    return 0;
}

}  // namespace pwc
