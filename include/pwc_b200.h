/* pwc_b200.h -- C ABI of libpwc_b200.so: the sm_100a compute path behind the Python call
 * surface of daigo0927/pwcnet (PWCDCNet.__call__, modules.py callables, losses.py).
 *
 * The reference has no FFI layer of its own (it is pure TensorFlow-1.8 Python); every entry
 * point below replaces the TF graph ops that one reference function builds, and cites that
 * function.  Conventions for every call:
 *   - all pointers are DEVICE pointers to float32 unless said otherwise; tensors are NHWC;
 *   - a "*_cs" argument is the channel stride (floats between consecutive pixels) of that
 *     tensor, so an op can read from / write into a channel slot of a wider concat buffer
 *     (this is how tf.concat, modules.py:262-264,305, is eliminated); pass C for a dense tensor;
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work: no allocation, no
 *     synchronisation, no hidden state; entry points are re-entrant;
 *   - return value 0 = success; >0 = cudaError_t; <0 = PWC_E_* argument error.  Nothing throws
 *     or aborts.  pwc_last_error() returns a thread-local message for the last failure.
 */
#ifndef PWC_B200_H_
#define PWC_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define PWC_ABI_VERSION 1

#define PWC_E_BADARG   (-1)  /* null pointer, non-positive dim, unsupported combination      */
#define PWC_E_ALIGN    (-2)  /* pointer / stride alignment required by the kernel not met    */
#define PWC_E_NOTBUILT (-3)  /* entry point compiled out                                     */

int pwc_version(void);
const char* pwc_last_error(void);

/* CRC-32C (Castagnoli) of n host bytes, continuing from `crc` (0 to start).  HOST code (slice-by-8): the checksum
 * TensorFlow's checkpoint bundles carry per tensor and per table block (tf.train.Saver.save, train.py:95,166);
 * used by pwcnet_b200/checkpoint.py:save_checkpoint. */
unsigned int pwc_crc32c(const void* data, long long n, unsigned int crc);

/* CostVolumeLayer.__call__ -> get_cost x (2r+1)^2  (modules.py:164-204).
 * out[b,y,x,(v+r)*(2r+1)+(h+r)] = leaky_alpha( (1/C) sum_c f0[b,y,x,c] * f1[b,y+v,x+h,c] ),
 * zero outside the image.  If f0_copy != NULL the f0 tile staged on chip is also written to
 * f0_copy (channel stride f0_copy_cs) -- the estimator's concat slot (modules.py:262).
 * Requires C % 4 == 0, 16-byte aligned f0/f1 and f0_cs % 4 == 0, f1_cs % 4 == 0. */
int pwc_cost_volume_fwd(const float* f0, int f0_cs, const float* f1, int f1_cs,
                        float* out, int out_cs, float* f0_copy, int f0_copy_cs,
                        int B, int H, int W, int C, int search_range, float alpha, void* stream);

/* Fused WarpingLayer + CostVolumeLayer (model.py:109-112 -> modules.py:99-137,189-204):
 * f1 is warped by flow*flow_scale on chip while the halo tile is staged; the warped features
 * never reach HBM.  warp_type 0 = bilinear (modules.py:99-137), 1 = nearest (modules.py:83-97). */
int pwc_warp_cost_volume_fwd(const float* f0, int f0_cs, const float* f1, int f1_cs,
                             const float* flow, int flow_cs, float flow_scale, int warp_type,
                             float* out, int out_cs, float* f0_copy, int f0_copy_cs,
                             int B, int H, int W, int C, int search_range, float alpha, void* stream);

/* WarpingLayer.__call__ (modules.py:139-154): bilinear_warp (99-137) / nearest_warp (83-97).
 * flow channel 0 = x, channel 1 = y, in pixels, multiplied by flow_scale first (model.py:109). */
int pwc_warp_fwd(const float* x, int x_cs, const float* flow, int flow_cs, float flow_scale,
                 int warp_type, float* out, int out_cs, int B, int H, int W, int C, void* stream);

/* tf.layers.Conv2D(filters,(3,3),strides,'same',dilation_rate) + bias [+ residual] [+ leaky]
 * (modules.py:62-67, 266-277, 306-326).  w is HWIO (3,3,Cin,Cout) exactly as in the reference
 * checkpoint; TF asymmetric SAME padding.  alpha = leaky slope (1.0f = no activation).
 * residual (may be NULL, channel stride res_cs) is added AFTER the activation-free head
 * (modules.py:275-277, 326).  CUDA-core fp32 path: any Cin/Cout/stride/dilation. */
int pwc_conv3x3_fwd(const float* x, int x_cs, const float* w_hwio, const float* bias,
                    const float* residual, int res_cs, float* y, int y_cs,
                    int B, int H, int W, int Cin, int Cout, int stride, int dilation,
                    float alpha, void* stream);

/* Same function on the tcgen05 tensor cores (implicit GEMM, TMA-staged taps, TMEM accumulators).
 * w_packed: [9][Cout_pad][Cin_pad] fp32 (tap-major, K contiguous) produced by
 * pwc_conv3x3_pack_weights; n_split = 1 (TF32) or 3 (3xTF32 error-compensated, fp32-class).
 * Requires stride 1 or 2, Cin == 16 or Cin >= 32, x 16-byte aligned, x_cs % 4 == 0, Cout % 16 == 0,
 * Cout <= 256. */
int pwc_conv3x3_tc_fwd(const float* x, int x_cs, const float* w_packed, const float* bias,
                       float* y, int y_cs, int B, int H, int W, int Cin, int Cout, int stride, int dilation,
                       float alpha, int n_split, void* stream);
/* bytes needed for w_packed (hi and lo planes) */
long long pwc_conv3x3_packed_bytes(int Cin, int Cout);
int pwc_conv3x3_pack_weights(const float* w_hwio, float* w_packed, int Cin, int Cout, void* stream);

/* Same function with the "3 x fp16, scaled residual" split (x = h + l * 2^-11, h/l fp16; tcgen05 kind::f16,
 * fp32 accumulation in TMEM): fp32-class accuracy at half the MMA instructions of 3xTF32.  Requires
 * |x|, |w| < 65504.  w_packed: [2][9][Cout][Cin_pad32] fp16 from pwc_conv3x3_pack_weights_f16.
 * Requires stride 1 or 2, Cin >= 16, Cout % 16 == 0, Cout <= 256, x 16-byte aligned, x_cs % 4 == 0. */
int pwc_conv3x3_tc_f16_fwd(const float* x, int x_cs, const void* w_packed, const float* bias,
                           float* y, int y_cs, int B, int H, int W, int Cin, int Cout, int stride, int dilation,
                           float alpha, void* stream);
/* Flow heads on the tensor cores (modules.py:274-277, 325-326): the same conv with the MMA N padded to Cout_pad
 * (multiple of 16; w_packed = pack of the kernel zero-padded to Cout_pad output channels, bias_pad zero-padded), only
 * the first Cout channels are stored, `residual` (may be NULL) is added after the activation.  stride 1. */
int pwc_conv3x3_tc_f16_head(const float* x, int x_cs, const void* w_packed, const float* bias_pad,
                            const float* residual, int res_cs, float* y, int y_cs, int B, int H, int W, int Cin,
                            int Cout, int Cout_pad, int dilation, float alpha, void* stream);
long long pwc_conv3x3_packed_bytes_f16(int Cin, int Cout);
int pwc_conv3x3_pack_weights_f16(const float* w_hwio, void* w_packed, int Cin, int Cout, void* stream);

/* Batched pack: `jobs` is a DEVICE array of n_jobs x 8 int64 {w, out, K, N, mode, a, b, c}; one launch for all layers.
 * mode 0: out = pack_weights_f16 of the (3,3,K,a) HWIO kernel w, output channels zero-padded to N (a <= N);
 * mode 1: out = pack_weights_f16(rot_weights(w, ci_begin=b, ci_count=c, ci_pad=N)) for the (3,3,a,K) HWIO kernel w,
 *         i.e. the stride-1 dgrad kernel (Conv2DBackpropInput, train.py:89) of input channels [b, b+c).
 * `out` sizes as pwc_conv3x3_packed_bytes_f16(K, N). */
int pwc_conv3x3_pack_weights_f16_batched(const long long* jobs, int n_jobs, void* stream);

/* tf.image.resize_bilinear(x,(OH,OW)) with align_corners=False as in TF 1.8 (no half-pixel
 * offset; modules.py:283-284, model.py:127), result multiplied by `mul` (model.py:127 "*20."). */
int pwc_resize_bilinear_fwd(const float* x, int x_cs, float* y, int y_cs, int B, int H, int W, int C,
                            int OH, int OW, float mul, void* stream);

/* losses.py:4-8,15-31: one pyramid level of multiscale_loss (ord 2 = L2loss, ord 1 = L1loss).
 * acc[0] += weight * (1/B) * sum_{b,y,x} || gt[b, y*H/h, x*W/w, :]/gt_div - fs[b,y,x,:] ||_ord
 * (resize_nearest_neighbor of the scaled ground truth, losses.py:20,27).  acc is a device float. */
int pwc_lploss_level_fwd(const float* gt, int H, int W, const float* fs, int fs_cs, int h, int w,
                         int B, float gt_div, float weight, int ord, float* acc, void* stream);
/* losses.py:11-13 EPE: acc[0] += (1/(B*H*W)) * sum || gt - flows ||_2 */
int pwc_epe_fwd(const float* gt, const float* flows, int B, int H, int W, float* acc, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Backward pass / training step (train.py:65-92; the node groups TF-1.8 autodiff built).
 * Gradient tensors have the shape and channel-stride conventions of the activations they belong to.
 * ------------------------------------------------------------------------------------------------ */

/* Conv2DBackpropInput of pwc_conv3x3_fwd.  H, W, Cin describe dx (the conv INPUT), dy is (B,OH,OW,Cout) with
 * the SAME-padding output size.  new = sum_{tap,co} dy * w;  if mask != NULL new *= (mask > 0 ? 1 : mask_alpha)
 * (LeakyReluGrad of the layer that produced the conv input; mask = that layer's output);
 * dx = accumulate ? dx + new : new. */
int pwc_conv3x3_dgrad(const float* dy, int dy_cs, const float* w_hwio, float* dx, int dx_cs,
                      const float* mask, int mask_cs, float mask_alpha, int accumulate,
                      int B, int H, int W, int Cin, int Cout, int stride, int dilation, void* stream);

/* Conv2DBackpropFilter + BiasAddGrad: dw[tap, map(ci), co] += sum x * dy and db[co] += sum dy (db may be NULL).
 * cin_map (device int32[Cin], may be NULL) maps the channel order of x (a concat buffer in internal order,
 * DESIGN.md section 2) to the input-channel index of the reference kernel (-1 = padding channel, skipped);
 * dw_cin = number of input channels of dw.  Accumulates with atomics: zero dw/db before the first call. */
int pwc_conv3x3_wgrad(const float* x, int x_cs, const float* dy, int dy_cs, float* dw, float* db,
                      const int* cin_map, int dw_cin, int B, int H, int W, int Cin, int Cout,
                      int stride, int dilation, void* stream);

/* LeakyReluGrad in place: g *= (y > 0 ? 1 : alpha), y = the activation OUTPUT (modules.py:63). */
int pwc_leaky_bwd(float* g, int g_cs, const float* y, int y_cs, long long n_pix, int C, float alpha, void* stream);

/* dst += scale * src over n_pix pixels x C channels (gradient fan-in of residual adds and concat slots). */
int pwc_add_strided(float* dst, int dst_cs, const float* src, int src_cs, long long n_pix, int C, float scale,
                    void* stream);

/* Zero insertion for the stride-2 dgrad (gradient of conv2d(strides=2), modules.py:60): out (B,H,W,C) dense,
 * out[b, 2y+oy, 2x+ox, :] = dy[b,y,x,:], zeros elsewhere; oy = 1 - pad_top, ox = 1 - pad_left of the SAME padding.
 * The stride-1 dgrad of `out` (pwc_conv3x3_tc_f16_dgrad) then equals the stride-2 dgrad of dy. */
int pwc_dilate2(const float* dy, int dy_cs, float* out, int B, int OH, int OW, int C, int H, int W, int oy, int ox,
                void* stream);

/* Gradient of pwc_cost_volume_fwd (modules.py:164-204).  g = gradient w.r.t. the cost volume OUTPUT, cv = that
 * output (for the leaky slope).  df0 += d/df0 (+ g_f0slot if not NULL: the gradient that arrived through the
 * f0 copy in the concat buffer); df1 = d/df1 (accumulate_f1 != 0: +=). */
int pwc_cost_volume_bwd(const float* g, int g_cs, const float* cv, int cv_cs, const float* f0, int f0_cs,
                        const float* f1, int f1_cs, const float* g_f0slot, int gs_cs,
                        float* df0, int df0_cs, float* df1, int df1_cs, int accumulate_f1,
                        int B, int H, int W, int C, int search_range, float alpha, void* stream);

/* Gradient of pwc_warp_fwd (modules.py:83-137): dx += scatter of the weighted taps (atomics; zero/initialise
 * dx first); dflow (may be NULL) += flow_scale * d/d(flow*flow_scale), through the bilinear weights only. */
int pwc_warp_bwd(const float* x, int x_cs, const float* flow, int flow_cs, float flow_scale, int warp_type,
                 const float* g, int g_cs, float* dx, int dx_cs, float* dflow, int dflow_cs,
                 int B, int H, int W, int C, void* stream);

/* ResizeBilinearGrad of pwc_resize_bilinear_fwd: dx (B,H,W,C) += adjoint applied to g (B,OH,OW,C) * mul. */
int pwc_resize_bilinear_bwd(const float* g, int g_cs, float* dx, int dx_cs, int B, int H, int W, int C,
                            int OH, int OW, float mul, void* stream);

/* Gradient of pwc_lploss_level_fwd w.r.t. fs: gfs (=|+=) weight/B * d||gt_s - fs||_ord / dfs. */
int pwc_lploss_level_bwd(const float* gt, int H, int W, const float* fs, int fs_cs, int h, int w, int B,
                         float gt_div, float weight, int ord, float* gfs, int gfs_cs, int accumulate,
                         void* stream);

/* train.py:74,89: g = grad*grad_scale + gamma*var (l2_loss regulariser); tf.train.AdamOptimizer update with
 * lr_t = lr*sqrt(1-beta2^t)/(1-beta1^t) read from the device float lr_t_dev (so a captured graph can be
 * replayed for every step); var -= lr_t * m / (sqrt(v) + eps). */
int pwc_adam_step(float* var, const float* grad, float* m, float* v, long long n, const float* lr_t_dev,
                  float beta1, float beta2, float eps, float gamma, float grad_scale, void* stream);

/* acc[0] += scale * sum x^2 (tf.nn.l2_loss with scale = 0.5, train.py:74). */
int pwc_sumsq(const float* x, long long n, float scale, float* acc, void* stream);

/* w_dst[tap, i, co] = perm[i] >= 0 ? w_src[tap, perm[i], co] : 0 -- internal-channel-order kernel of a concat
 * layer from the reference-order kernel (perm: device int32[cin_dst]). */
int pwc_permute_cin(const float* w_src, float* w_dst, const int* perm, int cin_dst, int cin_src, int cout,
                    void* stream);

/* w_rot[ky,kx,co,j] = w[2-ky,2-kx,ci_begin+j,co] (j < ci_count; zero for ci_count <= j < ci_pad): with it,
 * pwc_conv3x3_*_fwd(dy, w_rot, stride 1) IS the dgrad for input channels [ci_begin, ci_begin+ci_count). */
int pwc_conv3x3_rot_weights(const float* w_hwio, float* w_rot, int Cin, int Cout, int ci_begin, int ci_count,
                            int ci_pad, void* stream);

/* Conv2DBackpropInput of a stride-1 conv on tcgen05 (3 x fp16 split, fp32-class): dx[..., :Cdx] (=|+=)
 * conv(dy, w_rot) [* leaky'(mask)].  w_rot_packed = pwc_conv3x3_pack_weights_f16(pwc_conv3x3_rot_weights(w),
 * Cin = Cdy, Cout = Cdx_pad); Cdx_pad % 16 == 0, Cdx_pad <= 256, Cdy >= 16. */
int pwc_conv3x3_tc_f16_dgrad(const float* dy, int dy_cs, const void* w_rot_packed, float* dx, int dx_cs,
                             const float* mask, int mask_cs, float mask_alpha, int accumulate,
                             int B, int H, int W, int Cdy, int Cdx, int Cdx_pad, int dilation, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Cost volume from pre-split fp16 operands (the pipeline's fast path; cost_volume_tcs.cu).
 * A "split" tensor holds, per pixel and 32-channel slice, the 128-byte row [h: 32 x fp16 | l: 32 x fp16] with
 * h = fp16(x), l = fp16(x - h): the same bytes as the fp32 row, dense (B,H,W,2C) fp16.
 * ------------------------------------------------------------------------------------------------ */

/* The same stride-1 conv (tf.layers.Conv2D + leaky, modules.py:266-270, 306-323) inside a conv -> conv chain whose
 * intermediate tensors are kept as SPLIT rows: per pixel and 32-channel slice [h: 32 x fp16 | l: 32 x fp16 * 2^11] (the bytes
 * of 32 floats; the row layout of pwc_split_f16_fwd, whose l is unscaled), written by the producer's epilogue (y_split) and read by the consumer without its fp32 -> fp16 converter
 * pass (x_split).  Bit-identical to pwc_conv3x3_tc_f16_fwd on fp32 tensors.  Cout <= 128, dilation 1..16; x_cs / ys_cs count
 * halfs per pixel for split tensors; y and y_split may both be given (y may be NULL). */
int pwc_conv3x3_tc_f16_split_fwd(const void* x, int x_split, int x_cs, const void* w_packed, const float* bias,
                                 float* y, int y_cs, void* y_split, int ys_cs,
                                 int B, int H, int W, int Cin, int Cout, int dilation, float alpha, void* stream);

/* Stride-2 3x3 convolution + bias + leaky on the halo kernel (replaces modules.py:62-63, the first convolution of every
 * pyramid level: tf.layers.Conv2D(filters, (3,3), (2,2), 'same') + tf.nn.leaky_relu) as a 2x2 convolution over the
 * space-to-depth view of x, read in place through a 5-D tensor map.  x: [B,H,W,Cin] dense fp32, even H and W, Cin % 16 == 0,
 * Cout % 16 == 0, Cout <= 128; y / y_split as in pwc_conv3x3_tc_f16_split_fwd at [B,H/2,W/2].  w_packed =
 * pwc_conv3x3_pack_weights_f16(pwc_conv3x3_s2d_reindex(w_hwio), Cin' = 4*Cin, Cout). */
int pwc_conv3x3_s2d_reindex(const float* w_hwio, float* w_s2d, int Cin, int Cout, void* stream);
int pwc_conv3x3_s2_tc_f16_fwd(const float* x, const void* w_packed, const float* bias, float* y, int y_cs,
                              void* y_split, int ys_cs, int B, int H, int W, int Cin, int Cout, float alpha, void* stream);

/* First pyramid convolution (modules.py:62-63, l = 0): 3 -> 16 channels, 3x3, stride 2, SAME, + bias + leaky, exact fp32 on
 * the CUDA cores, reading either float32 RGB/255 images or (x_is_u8) the uint8 RGB bytes themselves through lut256
 * (= float32(float64(v)/255.0): the reference's `images/255.0`, test.py:31-33).  x: dense (B,H,W,3); y: (B,H/2,W/2,16)
 * with channel stride y_cs.  H even, W a multiple of 4.  Uploads the 448 weights to a module-global constant buffer in
 * stream order: calls with different weights on different streams must be serialised by the caller. */
int pwc_conv_first_fwd(const void* x, int x_is_u8, const float* lut256, const float* w_hwio, const float* bias,
                       float* y, int y_cs, int B, int H, int W, float alpha, void* stream);

/* *count += number of non-finite values in the pixel-strided tensor x (n_pix pixels, C channels, channel stride x_cs).
 * Range guard of the 3 x fp16 tensor-core path (operands must stay below 65504): an overflow anywhere upstream turns
 * the last pyramid flow into NaN; the host raises when the counter is non-zero.  No reference counterpart (TF computes
 * in fp32). */
int pwc_count_nonfinite(const float* x, int x_cs, int C, long long n_pix, int* count, void* stream);

/* uint8 RGB images -> float32 in [0,1]: y[i] = lut256[x[i]].  Replaces the reference's host-side
 * `np.array(images)/255.0` + float32 placeholder feed (test.py:31-33, train.py:122, test_continuous.py:49); with
 * lut256[v] = float32(float64(v)/255.0) the result is bit-identical to that feed.  x, y 16-byte aligned. */
int pwc_u8_to_f32_fwd(const unsigned char* x, float* y, long long n, const float* lut256, void* stream);

/* scale * x (fp32, channel stride x_cs) -> split tensor `out`; if copy != NULL the UNSCALED x is also copied to `copy`
 * (fp32, channel stride copy_cs: the f0 slot of the estimator's concat buffer, modules.py:262).  C % 32 == 0.
 * The pipeline passes scale = 1/C for f0, so the mean over channels (modules.py:181) costs nothing later. */
int pwc_split_f16_fwd(const float* x, int x_cs, void* out, float* copy, int copy_cs, long long n_pix, int C,
                      float scale, void* stream);

/* pwc_warp_fwd (WarpingLayer, modules.py:83-154) writing its result as a split tensor.  C % 32 == 0. */
int pwc_warp_split_fwd(const float* x, int x_cs, const float* flow, int flow_cs, float flow_scale, int warp_type,
                       void* out, int B, int H, int W, int C, void* stream);

/* pwc_cost_volume_fwd for search_range 4 on tcgen05 from split operands f0s, f1s (modules.py:164-204):
 * out[b,y,x,(v+4)*9+(h+4)] = leaky_alpha(scale * sum_c f0s[b,y,x,c] f1s[b,y+v,x+h,c]); scale = 1/C, or 1 when f0s was
 * produced with scale 1/C.  fp32-class (3 x fp16 products, fp32 accumulation). */
int pwc_cost_volume_split_fwd(const void* f0s, const void* f1s, float* out, int out_cs,
                              int B, int H, int W, int C, float scale, float alpha, void* stream);

/* Same kernel writing the HEAD of a concat-buffer pixel row in whole 32-byte sectors: words [0,81) = cost volume,
 * [81,83) = tail[b,y,x,0:2] (the up-sampled flow the estimator concatenates next, modules.py:262-264; zeros when tail is
 * NULL), [83,88) = zeros (padding channels whose weights are zero).  head must be 88; out 32-byte aligned, out_cs a
 * multiple of 8 floats.  Measured on B200 (profiles/r02_store_bw_bench.log, r02_cv_quad.log): partial-sector 324-byte runs
 * cost this kernel 30 %; 352-byte runs at a 608-byte pitch write at 4.6 TB/s, contiguous lines at 6.6 TB/s. */
int pwc_cost_volume_split_slot_fwd(const void* f0s, const void* f1s, float* out, int out_cs, int head, const float* tail,
                                   int B, int H, int W, int C, float scale, float alpha, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Conv2DBackpropFilter on tcgen05 (wgrad_tc.cu).  Operands are first transposed to channel-major fp16 planes
 * (h = fp16(v), l = fp16((v - h) * 2^11), each plane (B, C, H, OWp), OWp = OW rounded up to 8):
 *   n_shift = 1 (dy):   out = [h | l] of x itself (OW = W); if db != NULL, db[c] += sum x[..., c] (BiasAddGrad);
 *   n_shift = 3 (conv input): out = three copies [h | l], copy kx sampled at the conv's OUTPUT columns,
 *                       copy_kx[b, c, y, ox] = x[b, y, ox*stride - pad_left + kx*dilation, c] (0 outside).
 * pwc_tsplit_bytes gives the size of `out` (OW = output width for n_shift = 3).  C % 4 == 0.
 * ------------------------------------------------------------------------------------------------ */
long long pwc_tsplit_bytes(int B, int H, int OW, int C, int n_shift);
int pwc_tsplit_f16(const float* x, int x_cs, void* out, int B, int H, int W, int C, int n_shift, int stride,
                   int dilation, float* db, void* stream);

/* dw[tap, map(ci), co] += sum x * dy from xT = pwc_tsplit_f16(conv input, n_shift 3, stride, dilation) and
 * dyT = pwc_tsplit_f16(output gradient, n_shift 1); H, W, Cin describe the conv input; same cin_map / dw_cin meaning
 * as pwc_conv3x3_wgrad.  stride 1 or 2, Cout % 16 == 0.  Accumulates with atomics: zero dw before the first call. */
int pwc_conv3x3_wgrad_tc(const void* xT, const void* dyT, float* dw, const int* cin_map, int dw_cin,
                         int B, int H, int W, int Cin, int Cout, int stride, int dilation, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PWC_B200_H_ */
