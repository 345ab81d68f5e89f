"""CPU oracle for the PWC-Net hot path of daigo0927/pwcnet  --  TEST INFRASTRUCTURE ONLY.

This file restates, op for op and in the reference's own op structure, what
`modules.py`, `model.py` and `losses.py` of the reference compute (TensorFlow 1.8
graph semantics) on torch-CPU float32 (float64 on request).  It exists so that the
CUDA path in `pwcnet_b200/` can be checked against it.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs
may import it; the product package never does.

PARITY STATUS: TensorFlow 1.8 is not installable here (no cp312 wheel, no network) and the reference
ships no tests or golden vectors, so this oracle cannot be pinned by running the reference program.
It is pinned, in decreasing strength, by
  (1) the reference's OWN serialized computation: the TF-1.8 GraphDef it saved next to its checkpoints
      (`model_250.ckpt.meta`, forward + loss graphs built from model.py / modules.py / losses.py) is
      executed node by node by `oracle/tf_graph_interp.py` (a TensorFlow-free interpreter of the 30 op
      types that occur); this oracle matches it to 3e-7 (glorot weights), 4e-5 on 14-px flows and 2e-6
      with the trained checkpoint; the outputs are the fixtures in tests/golden/ (oracle/make_golden.py);
  (2) an end-to-end semantic check with the reference's trained checkpoint (recovers a known synthetic
      translation) and
  (3) hand-derived known-answer tests for every TF-1.8 kernel semantic (asymmetric SAME padding, legacy
      bilinear resize, clamp-not-zero warp, cost-volume channel order) in tests/test_oracle.py.
What remains unpinned: the numpy/torch kernels inside the interpreter restate TF's op kernels (Conv2D,
ResizeBilinear, GatherNd, ...) from their documented semantics, they are not TF's binaries.
See DESIGN.md "Oracle and parity".

All tensors are NHWC.  Conv kernels are HWIO.  `W` is a dict name -> array keyed by
the reference's checkpoint variable names (pwcdcnet/<scope>/conv2d[_i]/{kernel,bias}).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

PYRAMID_FILTERS = [16, 32, 64, 96, 128, 192]   # modules.py:45
ESTIMATOR_FILTERS = [128, 128, 96, 64, 32]     # modules.py:234
SCALES = [None, 0.625, 1.25, 2.5, 5.0, 10.0, 20.0]  # model.py:93


def _t(x, dtype=torch.float32) -> Tensor:
    if isinstance(x, Tensor):
        return x.to(dtype)
    return torch.from_numpy(np.ascontiguousarray(x)).to(dtype)


# --------------------------------------------------------------------------- conv
def same_padding(size: int, k: int, stride: int, dilation: int) -> Tuple[int, int]:
    """TF 'SAME' padding (before, after) for one spatial dim.

    tf.layers.Conv2D(..., 'same') as used at modules.py:62-66, 267, 274, 306-325.
    out = ceil(in/s); total = max((out-1)*s + (k-1)*d + 1 - in, 0); before = total//2.
    """
    out = -(-size // stride)
    total = max((out - 1) * stride + (k - 1) * dilation + 1 - size, 0)
    before = total // 2
    return before, total - before


def conv2d_same(x: Tensor, kernel: Tensor, bias: Optional[Tensor], stride: int = 1,
                dilation: int = 1) -> Tensor:
    """tf.layers.Conv2D(filters,(3,3),strides,'same',dilation_rate) : NHWC x HWIO -> NHWC.

    Cross-correlation (no flip) + bias, modules.py:62.  Dilated convs appear in the
    saved graph as SpaceToBatchND -> VALID conv -> BatchToSpaceND, which is the same
    function as a dilated SAME conv.
    """
    kh, kw = kernel.shape[0], kernel.shape[1]
    pt, pb = same_padding(x.shape[1], kh, stride, dilation)
    pl, pr = same_padding(x.shape[2], kw, stride, dilation)
    xn = x.permute(0, 3, 1, 2)
    xn = F.pad(xn, (pl, pr, pt, pb))
    w = kernel.permute(3, 2, 0, 1)  # HWIO -> OIHW
    y = F.conv2d(xn, w, bias, stride=stride, padding=0, dilation=dilation)
    return y.permute(0, 2, 3, 1).contiguous()


def leaky_relu(x: Tensor, alpha: float = 0.1) -> Tensor:
    """tf.nn.leaky_relu = max(alpha*x, x) (modules.py:63)."""
    return torch.maximum(alpha * x, x)


def _conv(W, scope: str, idx: int, x: Tensor, stride=1, dilation=1) -> Tensor:
    name = f"{scope}/conv2d" + (f"_{idx}" if idx else "")
    k = _t(W[name + "/kernel"], x.dtype)
    b = _t(W[name + "/bias"], x.dtype)
    return conv2d_same(x, k, b, stride, dilation)


# ------------------------------------------------------------------------ pyramid
def pyramid_extractor(images: Tensor, W, num_levels: int = 6, name="pwcdcnet/fp_extractor") -> List[Tensor]:
    """FeaturePyramidExtractor_custom.__call__ (modules.py:49-71): deep -> shallow list."""
    feats = []
    x = images
    for l in range(num_levels):
        x = leaky_relu(_conv(W, name, 3 * l + 0, x, stride=2), 0.1)
        x = leaky_relu(_conv(W, name, 3 * l + 1, x, stride=1), 0.1)
        x = leaky_relu(_conv(W, name, 3 * l + 2, x, stride=1), 0.1)
        feats.append(x)
    return feats[::-1]


# --------------------------------------------------------------------------- warp
def _gather_nd(x: Tensor, gy: Tensor, gx: Tensor) -> Tensor:
    """tf.gather_nd(x, stack([b, gy, gx])) with integer (long) gy, gx of shape (B,h,w)."""
    B = x.shape[0]
    b = torch.arange(B).view(B, 1, 1).expand_as(gy)
    return x[b, gy, gx]


def bilinear_warp(x: Tensor, flow: Tensor) -> Tensor:
    """bilinear_warp (modules.py:99-137).

    Tap indices are clamped to the image, the four weights come from the UN-clamped
    fractional part, so out-of-range samples replicate the border with full weight.
    flow[...,0] = x displacement, flow[...,1] = y displacement (modules.py:106).
    """
    B, h, w, _ = x.shape
    dt = x.dtype
    gy = torch.arange(h, dtype=dt).view(1, h, 1).expand(B, h, w)
    gx = torch.arange(w, dtype=dt).view(1, 1, w).expand(B, h, w)
    fx, fy = flow[..., 0], flow[..., 1]
    fx0 = torch.floor(fx); fx1 = fx0 + 1
    fy0 = torch.floor(fy); fy1 = fy0 + 1
    gy0 = torch.clamp(gy + fy0, 0.0, float(h - 1))
    gy1 = torch.clamp(gy + fy1, 0.0, float(h - 1))
    gx0 = torch.clamp(gx + fx0, 0.0, float(w - 1))
    gx1 = torch.clamp(gx + fx1, 0.0, float(w - 1))
    iy0, iy1, ix0, ix1 = (t.to(torch.int32).long() for t in (gy0, gy1, gx0, gx1))  # tf.cast -> int32
    x00 = _gather_nd(x, iy0, ix0)
    x01 = _gather_nd(x, iy0, ix1)
    x10 = _gather_nd(x, iy1, ix0)
    x11 = _gather_nd(x, iy1, ix1)
    c00 = ((fy1 - fy) * (fx1 - fx)).unsqueeze(3)
    c01 = ((fy1 - fy) * (fx - fx0)).unsqueeze(3)
    c10 = ((fy - fy0) * (fx1 - fx)).unsqueeze(3)
    c11 = ((fy - fy0) * (fx - fx0)).unsqueeze(3)
    return c00 * x00 + c01 * x01 + c10 * x10 + c11 * x11


def nearest_warp(x: Tensor, flow: Tensor) -> Tensor:
    """nearest_warp (modules.py:83-97): tf.cast(flow,int32) truncates toward zero."""
    B, h, w, _ = x.shape
    fi = torch.trunc(flow).to(torch.int64)
    gy = torch.arange(h).view(1, h, 1).expand(B, h, w)
    gx = torch.arange(w).view(1, 1, w).expand(B, h, w)
    wy = torch.clamp(gy + fi[..., 1], 0, h - 1)
    wx = torch.clamp(gx + fi[..., 0], 0, w - 1)
    return _gather_nd(x, wy, wx)


def warping_layer(x: Tensor, flow: Tensor, warp_type: str = "bilinear") -> Tensor:
    """WarpingLayer.__call__ (modules.py:144-154)."""
    assert warp_type in ["nearest", "bilinear"]
    return nearest_warp(x, flow) if warp_type == "nearest" else bilinear_warp(x, flow)


# -------------------------------------------------------------------- cost volume
def _pad2d(x: Tensor, vpad, hpad) -> Tensor:
    """pad2d (modules.py:158-159): tf.pad(x, [[0,0], vpad, hpad, [0,0]])."""
    return F.pad(x, (0, 0, hpad[0], hpad[1], vpad[0], vpad[1]))


def _crop2d(x: Tensor, vcrop, hcrop) -> Tensor:
    """crop2d (modules.py:161-162): Cropping2D([vcrop, hcrop])."""
    H, Wd = x.shape[1], x.shape[2]
    return x[:, vcrop[0]:H - vcrop[1], hcrop[0]:Wd - hcrop[1], :]


def get_cost(f0: Tensor, f1: Tensor, shift) -> Tensor:
    """get_cost (modules.py:164-181), same pad/pad/mul/crop/mean structure."""
    v, h = shift
    vt, vb, hl, hr = max(v, 0), abs(min(v, 0)), max(h, 0), abs(min(h, 0))
    f0p = _pad2d(f0, [vt, vb], [hl, hr])
    f1p = _pad2d(f1, [vb, vt], [hr, hl])
    cost_pad = f0p * f1p
    return torch.mean(_crop2d(cost_pad, [vt, vb], [hl, hr]), dim=3)


def cost_volume(f0: Tensor, f1w: Tensor, search_range: int = 4) -> Tensor:
    """CostVolumeLayer.__call__ (modules.py:189-204): v outer, h inner, stack, leaky 0.1."""
    cv = []
    for v in range(-search_range, search_range + 1):
        for h in range(-search_range, search_range + 1):
            cv.append(get_cost(f0, f1w, [v, h]))
    cv = torch.stack(cv, dim=3)
    return leaky_relu(cv, 0.1)


def cost_volume_closed_form(f0: Tensor, f1w: Tensor, search_range: int = 4) -> Tensor:
    """Same function via the closed form cv[b,y,x,(v+r)*(2r+1)+(h+r)] =
    leaky((1/C) sum_c f0[b,y,x,c] f1w[b,y+v,x+h,c]), zeros outside."""
    B, H, Wd, C = f0.shape
    r = search_range
    f1p = F.pad(f1w, (0, 0, r, r, r, r))
    out = []
    for v in range(-r, r + 1):
        for h in range(-r, r + 1):
            sl = f1p[:, r + v:r + v + H, r + h:r + h + Wd, :]
            out.append((f0 * sl).sum(dim=3) / C)
    return leaky_relu(torch.stack(out, dim=3), 0.1)


# ------------------------------------------------------------------------- resize
def resize_bilinear_legacy(x: Tensor, out_h: int, out_w: int) -> Tensor:
    """tf.image.resize_bilinear(x, (out_h,out_w)), align_corners=False, TF 1.8
    (modules.py:283-284, model.py:127): src = dst * in/out, NO half-pixel offset.

    Follows TF's kernel arithmetic: scale = in/out (float32), lower = floor(i*scale),
    upper = min(lower+1, in-1), lerp = i*scale - lower;
    top = tl + (tr-tl)*xl ; bot = bl + (br-bl)*xl ; out = top + (bot-top)*yl.
    """
    B, h, w, C = x.shape
    dt = x.dtype

    def interp(n_in, n_out):
        scale = np.float32(n_in) / np.float32(n_out)
        i = np.arange(n_out, dtype=np.float32)
        s = (i * scale).astype(np.float32)
        lo = np.floor(s).astype(np.int64)
        hi = np.minimum(lo + 1, n_in - 1)
        lerp = (s - lo.astype(np.float32)).astype(np.float32)
        return torch.from_numpy(lo), torch.from_numpy(hi), torch.from_numpy(lerp).to(dt)

    ylo, yhi, yl = interp(h, out_h)
    xlo, xhi, xl = interp(w, out_w)
    top_rows = x[:, ylo]            # (B,out_h,w,C)
    bot_rows = x[:, yhi]
    xl_ = xl.view(1, 1, out_w, 1)
    yl_ = yl.view(1, out_h, 1, 1)
    tl, tr = top_rows[:, :, xlo], top_rows[:, :, xhi]
    bl, br = bot_rows[:, :, xlo], bot_rows[:, :, xhi]
    top = tl + (tr - tl) * xl_
    bot = bl + (br - bl) * xl_
    return top + (bot - top) * yl_


def resize_nearest_legacy(x: Tensor, out_h: int, out_w: int) -> Tensor:
    """tf.image.resize_nearest_neighbor, align_corners=False (losses.py:27):
    src = min(floor(i * in/out), in-1)."""
    B, h, w, C = x.shape
    sy = np.float32(h) / np.float32(out_h)
    sx = np.float32(w) / np.float32(out_w)
    iy = np.minimum(np.floor(np.arange(out_h, dtype=np.float32) * sy).astype(np.int64), h - 1)
    ix = np.minimum(np.floor(np.arange(out_w, dtype=np.float32) * sx).astype(np.int64), w - 1)
    return x[:, torch.from_numpy(iy)][:, :, torch.from_numpy(ix)]


# ------------------------------------------------------------------ flow estimator
def flow_estimator(W, scope: str, cv: Tensor, features_0: Optional[Tensor], flows_up_prev: Optional[Tensor],
                   features_up_prev: Optional[Tensor], is_output: bool = False, use_dc: bool = False):
    """OpticalFlowEstimator_custom.__call__ (modules.py:260-285)."""
    features = cv
    for f in [features_0, flows_up_prev, features_up_prev]:
        if f is not None:
            features = torch.cat([features, f], dim=3)
    for i in range(len(ESTIMATOR_FILTERS)):
        conv = leaky_relu(_conv(W, scope, i, features), 0.1)
        features = torch.cat([conv, features], dim=3) if use_dc else conv
    flows = _conv(W, scope, len(ESTIMATOR_FILTERS), features)
    if flows_up_prev is not None:
        flows = flows + flows_up_prev
    if is_output:
        return flows, features
    h, w = flows.shape[1], flows.shape[2]
    flows_up = resize_bilinear_legacy(flows, 2 * h, 2 * w)
    features_up = resize_bilinear_legacy(features, 2 * h, 2 * w)
    return flows, flows_up, features_up


CONTEXT_DILATIONS = [1, 2, 4, 8, 16, 1, 1]  # modules.py:306-325


def context_network(W, flows: Tensor, features: Tensor, scope="pwcdcnet/context") -> Tensor:
    """ContextNetwork.__call__ (modules.py:304-326)."""
    x = torch.cat([flows, features], dim=3)
    for i, d in enumerate(CONTEXT_DILATIONS):
        x = _conv(W, scope, i, x, dilation=d)
        if i < 6:
            x = leaky_relu(x, 0.1)
    return flows + x


# -------------------------------------------------------------------------- model
def pwcdcnet_forward(W, images_0, images_1, num_levels: int = 6, search_range: int = 4,
                     warp_type: str = "bilinear", use_dc: bool = False, output_level: int = 4,
                     name: str = "pwcdcnet", dtype=torch.float32, return_intermediates: bool = False):
    """PWCDCNet.__call__ (model.py:95-134) -> (flows_final, flows_pyramid[, intermediates])."""
    assert output_level < num_levels
    im0, im1 = _t(images_0, dtype), _t(images_1, dtype)
    pyr0 = pyramid_extractor(im0, W, num_levels, f"{name}/fp_extractor")
    pyr1 = pyramid_extractor(im1, W, num_levels, f"{name}/fp_extractor")
    flows_pyramid = []
    inter = {"pyramid_0": pyr0, "pyramid_1": pyr1, "cv": [], "warped": []}
    flows_up, features_up = None, None
    for l, (f0, f1) in enumerate(zip(pyr0, pyr1)):
        if l == 0:
            f1w = f1
        else:
            f1w = warping_layer(f1, flows_up * SCALES[l], warp_type)
        cv = cost_volume(f0, f1w, search_range)
        inter["cv"].append(cv); inter["warped"].append(f1w)
        if l < output_level:
            flows, flows_up, features_up = flow_estimator(W, f"{name}/optflow_{l}", cv, f0, flows_up,
                                                          features_up, use_dc=use_dc)
        else:
            flows, features = flow_estimator(W, f"{name}/optflow_{l}", cv, f0, flows_up, features_up,
                                             is_output=True, use_dc=use_dc)
            flows = context_network(W, flows, features, f"{name}/context")
            flows_pyramid.append(flows)
            upscale = 2 ** (num_levels - output_level)
            h, w = flows.shape[1], flows.shape[2]
            flows_final = resize_bilinear_legacy(flows, h * upscale, w * upscale) * 20.0
            if return_intermediates:
                return flows_final, flows_pyramid, inter
            return flows_final, flows_pyramid
        flows_pyramid.append(flows)


# ------------------------------------------------------------------------- losses
def L1loss(x: Tensor, y: Tensor) -> Tensor:
    """losses.py:4-5."""
    return torch.mean(torch.sum(torch.sum(torch.abs(x - y), dim=3), dim=(1, 2)))


def L2loss(x: Tensor, y: Tensor) -> Tensor:
    """losses.py:7-8: mean_b( sum_{h,w} ||x-y||_2 )."""
    return torch.mean(torch.sum(torch.sqrt(torch.sum((x - y) ** 2, dim=3)), dim=(1, 2)))


def EPE(flows_gt: Tensor, flows: Tensor) -> Tensor:
    """losses.py:11-13."""
    return torch.mean(torch.sqrt(torch.sum((flows_gt - flows) ** 2, dim=3)))


def multiscale_loss(flows_gt: Tensor, flows_pyramid: Sequence[Tensor], weights: Sequence[float]) -> Tensor:
    """losses.py:15-31."""
    gt_scaled = flows_gt / 20.0
    loss = 0.0
    for weight, fs in zip(weights, flows_pyramid):
        h, w = fs.shape[1], fs.shape[2]
        gt_down = resize_nearest_legacy(gt_scaled, h, w)
        loss = loss + weight * L2loss(gt_down, fs)
    return loss


DEFAULT_LOSS_WEIGHTS = [0.32, 0.08, 0.02, 0.01, 0.005]  # train.py:220-222


def training_loss(W_t: Dict[str, Tensor], images_0, images_1, flows_gt, weights=DEFAULT_LOSS_WEIGHTS,
                  gamma: float = 4e-4, **kw):
    """train.py:65-77: multiscale loss + gamma * sum_v 0.5*||v||^2 over ALL model vars; EPE."""
    flows_final, pyr = pwcdcnet_forward(W_t, images_0, images_1, **kw)
    dt = flows_final.dtype
    gt = _t(flows_gt, dt)
    _loss = multiscale_loss(gt, pyr, weights)
    l2 = sum(0.5 * torch.sum(v.to(dt) ** 2) for v in W_t.values())
    return _loss + gamma * l2, EPE(gt, flows_final), flows_final, pyr


def piecewise_lr(step: int, lr: float = 1e-4) -> float:
    """train.py:82-85: tf.train.piecewise_constant(global_step, boundaries, values);
    value i applies for boundaries[i-1] < step <= boundaries[i]."""
    boundaries = [200000, 250000, 300000, 350000, 4000000]
    values = [lr / (2 ** i) for i in range(len(boundaries) + 1)]
    for b, v in zip(boundaries, values):
        if step <= b:
            return v
    return values[-1]


def adam_step_tf(var: np.ndarray, grad: np.ndarray, m: np.ndarray, v: np.ndarray, t: int, lr: float,
                 beta1=0.9, beta2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer update (train.py:89): lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; var -= lr_t * m / (sqrt(v) + eps)."""
    lr_t = np.float32(lr * math.sqrt(1 - beta2 ** t) / (1 - beta1 ** t))
    m = (beta1 * m + (1 - beta1) * grad).astype(np.float32)
    v = (beta2 * v + (1 - beta2) * grad * grad).astype(np.float32)
    var = (var - lr_t * m / (np.sqrt(v) + np.float32(eps))).astype(np.float32)
    return var, m, v


# ------------------------------------------------------------------------ weights
def layer_table(num_levels: int = 6, use_dc: bool = False, search_range: int = 4, name: str = "pwcdcnet",
                output_level: int = 4):
    """(variable scope name, Cin, Cout) for every pyramid / estimator conv of PWCDCNet in creation
    order (SURVEY 9.1 naming; shapes cross-checked against the checkpoints' .index).  Estimators
    above `output_level` are constructed (model.py:90-91) but never called, so TF never creates
    their variables."""
    rows = []
    cin = 3
    for l in range(num_levels):
        for j in range(3):
            idx = 3 * l + j
            rows.append((f"{name}/fp_extractor/conv2d" + (f"_{idx}" if idx else ""), cin, PYRAMID_FILTERS[l]))
            cin = PYRAMID_FILTERS[l]
    ncv = (2 * search_range + 1) ** 2
    feats_deep_first = PYRAMID_FILTERS[:num_levels][::-1]
    up_ch = None
    for l in range(output_level + 1):
        c = ncv + feats_deep_first[l] + (0 if l == 0 else 2 + up_ch)
        for i, f in enumerate(ESTIMATOR_FILTERS):
            rows.append((f"{name}/optflow_{l}/conv2d" + (f"_{i}" if i else ""), c, f))
            c = f + c if use_dc else f
        rows.append((f"{name}/optflow_{l}/conv2d_{len(ESTIMATOR_FILTERS)}", c, 2))
        up_ch = c
    return rows, up_ch


def glorot_weights(seed: int = 2, num_levels: int = 6, use_dc: bool = False, search_range: int = 4,
                   output_level: int = 4, name: str = "pwcdcnet", bias_scale: float = 0.0,
                   gain: float = 1.0) -> Dict[str, np.ndarray]:
    """Seeded glorot-uniform kernels (limit sqrt(6/(9Cin+9Cout))) and zero biases - the
    initialisers recorded in the reference GraphDef.  `gain`/`bias_scale`
    let tests make 'hot' weights whose flows are several pixels (exercises warp clamping)."""
    rng = np.random.default_rng(seed)
    rows, _ = layer_table(num_levels, use_dc, search_range, name, output_level)
    W = {}
    est_out = {}
    for scope, cin, cout in rows:
        lim = math.sqrt(6.0 / (9 * cin + 9 * cout)) * gain
        W[scope + "/kernel"] = rng.uniform(-lim, lim, size=(3, 3, cin, cout)).astype(np.float32)
        W[scope + "/bias"] = (rng.standard_normal(cout) * bias_scale).astype(np.float32)
        est_out[scope] = cin
    # context net: input = [flows(2), features(C_feat at output level)]
    feat_c = est_out[f"{name}/optflow_{output_level}/conv2d_{len(ESTIMATOR_FILTERS)}"]
    cin = 2 + feat_c
    for i, cout in enumerate([128, 128, 128, 96, 64, 32, 2]):
        scope = f"{name}/context/conv2d" + (f"_{i}" if i else "")
        lim = math.sqrt(6.0 / (9 * cin + 9 * cout)) * gain
        W[scope + "/kernel"] = rng.uniform(-lim, lim, size=(3, 3, cin, cout)).astype(np.float32)
        W[scope + "/bias"] = (rng.standard_normal(cout) * bias_scale).astype(np.float32)
        cin = cout
    return W


def synthetic_pair(B: int, H: int, W: int, seed: int = 0, shift: Optional[Tuple[int, int]] = None):
    """SURVEY 8(d) synthetic inputs: images rng.random((B,2,H,W,3)); optional shifted-texture
    pair (image_1 = image_0 rolled by (dx,dy)) so that the true flow is non-trivial."""
    rng = np.random.default_rng(seed)
    imgs = rng.random((B, 2, H, W, 3), dtype=np.float32)
    if shift is not None:
        dx, dy = shift
        imgs[:, 1] = np.roll(imgs[:, 0], (dy, dx), axis=(1, 2))
    return imgs[:, 0].copy(), imgs[:, 1].copy()


def synthetic_textured_pair(B: int, H: int, W: int, seed: int = 0, max_disp: float = 12.0):
    """SURVEY 8(d) 'shifted-texture pair': image_0 = smooth multi-octave random texture in [0,1], image_1 = image_0
    displaced by a smooth known field of up to +-max_disp pixels (bilinear resampling, numpy only).  Returns
    (image_0, image_1, flow) with flow (B,H,W,2) = the displacement image_1 was built from (ch0 = x, ch1 = y):
    image_1(y, x) = image_0(y - flow_y, x - flow_x), i.e. content moves by +flow from frame 0 to frame 1."""
    rng = np.random.default_rng(seed)
    tex = np.zeros((B, H, W, 3), np.float64)
    amp = 0.0
    for octave, cell in enumerate((64, 32, 16, 8, 4, 2)):
        gh, gw = -(-H // cell) + 2, -(-W // cell) + 2
        g = rng.random((B, gh, gw, 3))
        yy = (np.arange(H) + 0.5) / cell
        xx = (np.arange(W) + 0.5) / cell
        y0, x0 = np.floor(yy).astype(int), np.floor(xx).astype(int)
        fy, fx = (yy - y0)[None, :, None, None], (xx - x0)[None, None, :, None]
        a = g[:, y0][:, :, x0] * (1 - fx) + g[:, y0][:, :, x0 + 1] * fx
        b = g[:, y0 + 1][:, :, x0] * (1 - fx) + g[:, y0 + 1][:, :, x0 + 1] * fx
        wgt = 0.5 ** octave if octave < 4 else 0.5 ** 3
        tex += wgt * (a * (1 - fy) + b * fy)
        amp += wgt
    im0 = (tex / amp).astype(np.float32)
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    flow = np.zeros((B, H, W, 2), np.float32)
    im1 = np.empty_like(im0)
    for b in range(B):
        ph = rng.uniform(0, 2 * np.pi, 4)
        fx_ = max_disp * np.sin(2 * np.pi * ys / H + ph[0]) * np.cos(2 * np.pi * xs / W + ph[1])
        fy_ = 0.5 * max_disp * np.cos(2 * np.pi * ys / H + ph[2]) * np.sin(4 * np.pi * xs / W + ph[3])
        flow[b, ..., 0], flow[b, ..., 1] = fx_, fy_
        sy, sx = np.clip(ys - fy_, 0, H - 1), np.clip(xs - fx_, 0, W - 1)
        y0, x0 = np.floor(sy).astype(int), np.floor(sx).astype(int)
        y1, x1 = np.minimum(y0 + 1, H - 1), np.minimum(x0 + 1, W - 1)
        wy, wx = (sy - y0)[..., None], (sx - x0)[..., None]
        src = im0[b].astype(np.float64)
        im1[b] = ((src[y0, x0] * (1 - wx) + src[y0, x1] * wx) * (1 - wy) +
                  (src[y1, x0] * (1 - wx) + src[y1, x1] * wx) * wy).astype(np.float32)
    return im0, im1, flow


# ------------------------------------------------------------------ PWCNet (model.py:6-71), repaired
# The reference's `PWCNet` cannot be instantiated.  This restates its INTENDED design with the minimal
# repairs listed in SURVEY 2.4 -- parity unpinned against the reference (nothing executable, no checkpoint, no GraphDef):
#   model.py:19  OpticalFlowEstimator(self.batch_norm, ...)  -> OpticalFlowEstimator(name=...) (batch_norm False, modules.py:210)
#   model.py:23  self.context (never set)                    -> 'final' (one ContextNetwork applied at output_level)
#   model.py:50  of_estimators[l](feature_0, cost, flow)      -> kept as written: positional (cost, x, flow) = (feature_0,
#                                                               cost, flow), i.e. concat [feature_0, cost, flow] (modules.py:216)
#   model.py:56  context_net(feature, flow)                  -> context_net(flow, feature) (ContextNetwork.__call__(flows, features))
def pwcnet_layer_table(num_levels: int = 6, search_range: int = 4, output_level: int = 4, name: str = "pwcnet"):
    rows, cin = [], 3
    for l in range(num_levels):                      # FeaturePyramidExtractor (modules.py:19-39): two convs per level
        for j in range(2):
            idx = 2 * l + j
            rows.append((f"{name}/fp_extractor/conv2d" + (f"_{idx}" if idx else ""), cin, PYRAMID_FILTERS[l]))
            cin = PYRAMID_FILTERS[l]
    ncv = (2 * search_range + 1) ** 2
    deep_first = PYRAMID_FILTERS[:num_levels][::-1]
    for l in range(output_level + 1):                # OpticalFlowEstimator (modules.py:208-224)
        c = deep_first[l] + ncv + 2
        for i, f in enumerate(ESTIMATOR_FILTERS + [2]):
            rows.append((f"{name}/optflow_{l}/conv2d" + (f"_{i}" if i else ""), c, f))
            c = f
    cin = 2 + ESTIMATOR_FILTERS[-1]
    for i, cout in enumerate([128, 128, 128, 96, 64, 32, 2]):
        rows.append((f"{name}/context/conv2d" + (f"_{i}" if i else ""), cin, cout))
        cin = cout
    return rows


def pwcnet_glorot_weights(seed: int = 2, gain: float = 1.0, bias_scale: float = 0.0, **kw) -> Dict[str, np.ndarray]:
    rng = np.random.default_rng(seed)
    W = {}
    for scope, cin, cout in pwcnet_layer_table(**kw):
        lim = math.sqrt(6.0 / (9 * cin + 9 * cout)) * gain
        W[scope + "/kernel"] = rng.uniform(-lim, lim, size=(3, 3, cin, cout)).astype(np.float32)
        W[scope + "/bias"] = (rng.standard_normal(cout) * bias_scale).astype(np.float32)
    return W


def pwcnet_forward(W, images_0, images_1, num_levels: int = 6, search_range: int = 4, warp_type: str = "bilinear",
                   output_level: int = 4, name: str = "pwcnet", dtype=torch.float32):
    """PWCNet.__call__ (model.py:30-67) with the repairs above -> (finalflow, flows, pyramid_0)."""
    def pyramid(x):
        feats = []
        for l in range(num_levels):                  # modules.py:31-36
            x = leaky_relu(_conv(W, f"{name}/fp_extractor", 2 * l, x, stride=2), 0.1)
            x = leaky_relu(_conv(W, f"{name}/fp_extractor", 2 * l + 1, x, stride=1), 0.1)
            feats.append(x)
        return feats[::-1]
    pyr0, pyr1 = pyramid(_t(images_0, dtype)), pyramid(_t(images_1, dtype))
    flows, flow = [], None
    for l, (f0, f1) in enumerate(zip(pyr0, pyr1)):
        b, h, w, _ = f0.shape
        flow = torch.zeros((b, h, w, 2), dtype=dtype) if l == 0 else resize_bilinear_legacy(flow, h, w) * 2   # model.py:42-45
        f1w = warping_layer(f1, flow, warp_type)                                                              # model.py:48
        cost = cost_volume(f0, f1w, search_range)
        x = torch.cat([f0, cost, flow], dim=3)                                                                 # modules.py:216 via model.py:50
        for i in range(len(ESTIMATOR_FILTERS)):
            x = leaky_relu(_conv(W, f"{name}/optflow_{l}", i, x), 0.2)                                        # _conv_block, modules.py:7-15
        feature = x
        flow = _conv(W, f"{name}/optflow_{l}", len(ESTIMATOR_FILTERS), feature)                               # modules.py:222
        if l == output_level:
            flow = context_network(W, flow, feature, f"{name}/context")                                        # model.py:55-56 (repaired)
        flows.append(flow)
        if l == output_level:
            upscale = 2 ** (num_levels - output_level)
            return resize_bilinear_legacy(flow, h * upscale, w * upscale) * upscale, flows, pyr0               # model.py:62-64,67
