"""Backward operators (csrc/backward.cu) over pixel-strided NHWC float32 CUDA views.

Each function launches one hand-written kernel through the C ABI; gradients are written into / accumulated
into caller-owned buffers that mirror the activation buffers (same shapes and channel strides), so the fan-in
of concat slots and residual adds needs no extra tensors.  No autograd, no fallbacks."""
from __future__ import annotations

import torch

from ._abi import check, lib
from .ops import WARP_TYPES, _nhwc, _same_out, _stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def conv3x3_dgrad(dy, kernel, dx, stride: int = 1, dilation: int = 1, mask=None, mask_alpha: float = 0.1,
                  accumulate: bool = False):
    """dx (=|+=) Conv2DBackpropInput(dy, kernel) [* leaky'(mask)].  kernel: HWIO of the forward conv."""
    B, H, W, Cin, dx_cs = _nhwc(dx, "dx")
    Bo, OH, OW, Cout, dy_cs = _nhwc(dy, "dy")
    if tuple(kernel.shape) != (3, 3, Cin, Cout) or not kernel.is_contiguous():
        raise ValueError(f"conv3x3_dgrad: kernel {tuple(kernel.shape)} does not match dx/dy channels {(Cin, Cout)}")
    if (Bo, OH, OW) != (B, _same_out(H, stride), _same_out(W, stride)):
        raise ValueError("conv3x3_dgrad: dy spatial shape is not the SAME-padding output of dx")
    m_cs = 0
    if mask is not None:
        Bm, Hm, Wm, Cm, m_cs = _nhwc(mask, "mask")
        if (Bm, Hm, Wm, Cm) != (B, H, W, Cin):
            raise ValueError("conv3x3_dgrad: mask shape mismatch")
    check(lib().pwc_conv3x3_dgrad(dy.data_ptr(), dy_cs, kernel.data_ptr(), dx.data_ptr(), dx_cs, _ptr(mask), m_cs,
                                  float(mask_alpha), int(accumulate), B, H, W, Cin, Cout, stride, dilation, _stream()),
          "pwc_conv3x3_dgrad")
    return dx


def conv3x3_wgrad(x, dy, dw, db=None, stride: int = 1, dilation: int = 1, cin_map=None):
    """dw += Conv2DBackpropFilter(x, dy), db += BiasAddGrad(dy).  dw: (3,3,Cin_ref,Cout) contiguous."""
    B, H, W, Cin, x_cs = _nhwc(x, "x")
    Bo, OH, OW, Cout, dy_cs = _nhwc(dy, "dy")
    if (Bo, OH, OW) != (B, _same_out(H, stride), _same_out(W, stride)):
        raise ValueError("conv3x3_wgrad: dy spatial shape is not the SAME-padding output of x")
    if dw.dim() != 4 or dw.shape[:2] != (3, 3) or dw.shape[3] != Cout or not dw.is_contiguous():
        raise ValueError(f"conv3x3_wgrad: dw must be contiguous (3,3,Cin,{Cout}), got {tuple(dw.shape)}")
    if cin_map is None and dw.shape[2] != Cin:
        raise ValueError("conv3x3_wgrad: dw input channels differ from x (pass cin_map for concat layers)")
    if cin_map is not None and (cin_map.dtype != torch.int32 or cin_map.numel() != Cin or not cin_map.is_cuda):
        raise ValueError("conv3x3_wgrad: cin_map must be a CUDA int32 tensor with one entry per channel of x")
    if db is not None and (db.shape != (Cout,) or not db.is_contiguous()):
        raise ValueError("conv3x3_wgrad: db must be contiguous (Cout,)")
    check(lib().pwc_conv3x3_wgrad(x.data_ptr(), x_cs, dy.data_ptr(), dy_cs, dw.data_ptr(), _ptr(db), _ptr(cin_map),
                                  dw.shape[2], B, H, W, Cin, Cout, stride, dilation, _stream()), "pwc_conv3x3_wgrad")
    return dw


def leaky_bwd(g, y, alpha: float = 0.1):
    """g *= leaky'(y) in place (y = the activation output)."""
    B, H, W, C, g_cs = _nhwc(g, "g")
    B2, H2, W2, C2, y_cs = _nhwc(y, "y")
    if (B, H, W, C) != (B2, H2, W2, C2):
        raise ValueError("leaky_bwd: shape mismatch")
    check(lib().pwc_leaky_bwd(g.data_ptr(), g_cs, y.data_ptr(), y_cs, B * H * W, C, float(alpha), _stream()), "pwc_leaky_bwd")
    return g


def add_(dst, src, scale: float = 1.0):
    """dst += scale * src (both pixel-strided NHWC views of the same shape)."""
    B, H, W, C, d_cs = _nhwc(dst, "dst")
    B2, H2, W2, C2, s_cs = _nhwc(src, "src")
    if (B, H, W, C) != (B2, H2, W2, C2):
        raise ValueError("add_: shape mismatch")
    check(lib().pwc_add_strided(dst.data_ptr(), d_cs, src.data_ptr(), s_cs, B * H * W, C, float(scale), _stream()),
          "pwc_add_strided")
    return dst


def dilate2(dy, out, oy: int, ox: int):
    """Zero insertion: out[b, 2y+oy, 2x+ox] = dy[b, y, x], zeros elsewhere (out dense NHWC with dy's channel count)."""
    B, OH, OW, C, dy_cs = _nhwc(dy, "dy")
    Bo, H, W, Co, o_cs = _nhwc(out, "out")
    if Bo != B or Co != C or o_cs != C:
        raise ValueError("dilate2: out must be dense NHWC with dy's batch and channels")
    check(lib().pwc_dilate2(dy.data_ptr(), dy_cs, out.data_ptr(), B, OH, OW, C, H, W, int(oy), int(ox), _stream()), "pwc_dilate2")
    return out


def cost_volume_bwd(g, cv, f0, f1, df0, df1, g_f0slot=None, accumulate_f1: bool = False, search_range: int = 4,
                    alpha: float = 0.1):
    """df0 += dCV/df0 (+ g_f0slot), df1 (=|+=) dCV/df1 for cv = cost_volume(f0, f1)."""
    B, H, W, C, f0_cs = _nhwc(f0, "f0")
    _, _, _, _, f1_cs = _nhwc(f1, "f1")
    nd = (2 * search_range + 1) ** 2
    Bg, Hg, Wg, Cg, g_cs = _nhwc(g, "g")
    _, _, _, Cc, cv_cs = _nhwc(cv, "cv")
    if (Bg, Hg, Wg, Cg) != (B, H, W, nd) or Cc != nd or f1.shape != f0.shape or df0.shape != f0.shape or df1.shape != f0.shape:
        raise ValueError("cost_volume_bwd: shape mismatch")
    _, _, _, _, df0_cs = _nhwc(df0, "df0")
    _, _, _, _, df1_cs = _nhwc(df1, "df1")
    gs_cs = 0
    if g_f0slot is not None:
        if g_f0slot.shape != f0.shape:
            raise ValueError("cost_volume_bwd: g_f0slot shape mismatch")
        gs_cs = _nhwc(g_f0slot, "g_f0slot")[4]
    check(lib().pwc_cost_volume_bwd(g.data_ptr(), g_cs, cv.data_ptr(), cv_cs, f0.data_ptr(), f0_cs, f1.data_ptr(), f1_cs,
                                    _ptr(g_f0slot), gs_cs, df0.data_ptr(), df0_cs, df1.data_ptr(), df1_cs,
                                    int(accumulate_f1), B, H, W, C, search_range, float(alpha), _stream()),
          "pwc_cost_volume_bwd")
    return df0, df1


def warp_bwd(x, flow, g, dx, dflow=None, flow_scale: float = 1.0, warp_type: str = "bilinear"):
    """dx += scatter(g) ; dflow += flow_scale * d warp / d(flow * flow_scale)."""
    B, H, W, C, x_cs = _nhwc(x, "x")
    Bf, Hf, Wf, Cf, fl_cs = _nhwc(flow, "flow")
    if (Bf, Hf, Wf, Cf) != (B, H, W, 2) or g.shape != x.shape or dx.shape != x.shape:
        raise ValueError("warp_bwd: shape mismatch")
    g_cs = _nhwc(g, "g")[4]
    dx_cs = _nhwc(dx, "dx")[4]
    df_cs = 0
    if dflow is not None:
        if dflow.shape != flow.shape:
            raise ValueError("warp_bwd: dflow shape mismatch")
        df_cs = _nhwc(dflow, "dflow")[4]
    check(lib().pwc_warp_bwd(x.data_ptr(), x_cs, flow.data_ptr(), fl_cs, float(flow_scale), WARP_TYPES[warp_type],
                             g.data_ptr(), g_cs, dx.data_ptr(), dx_cs, _ptr(dflow), df_cs, B, H, W, C, _stream()),
          "pwc_warp_bwd")
    return dx


def resize_bilinear_bwd(g, dx, mul: float = 1.0):
    """dx += adjoint of the legacy bilinear resize applied to g * mul."""
    B, OH, OW, C, g_cs = _nhwc(g, "g")
    B2, H, W, C2, dx_cs = _nhwc(dx, "dx")
    if (B, C) != (B2, C2):
        raise ValueError("resize_bilinear_bwd: shape mismatch")
    check(lib().pwc_resize_bilinear_bwd(g.data_ptr(), g_cs, dx.data_ptr(), dx_cs, B, H, W, C, OH, OW, float(mul), _stream()),
          "pwc_resize_bilinear_bwd")
    return dx


def lploss_level_bwd(flows_gt, fs, weight: float, gfs, gt_div: float = 20.0, ord: int = 2, accumulate: bool = False):
    """gfs (=|+=) d[ weight * L{ord}loss(resize_nearest(flows_gt/gt_div), fs) ] / dfs."""
    B, H, W, C, g_cs = _nhwc(flows_gt, "flows_gt")
    Bf, h, w, Cf, f_cs = _nhwc(fs, "flows")
    if C != 2 or Cf != 2 or B != Bf or g_cs != 2 or gfs.shape != fs.shape:
        raise ValueError("lploss_level_bwd: shape mismatch")
    gf_cs = _nhwc(gfs, "gfs")[4]
    check(lib().pwc_lploss_level_bwd(flows_gt.data_ptr(), H, W, fs.data_ptr(), f_cs, h, w, B, float(gt_div), float(weight),
                                     int(ord), gfs.data_ptr(), gf_cs, int(accumulate), _stream()), "pwc_lploss_level_bwd")
    return gfs


def adam_step(var, grad, m, v, lr_t_dev, beta1=0.9, beta2=0.999, eps=1e-8, gamma=0.0, grad_scale=1.0):
    """Flat fp32 buffers; lr_t_dev: 1-element CUDA float tensor holding lr*sqrt(1-b2^t)/(1-b1^t)."""
    n = var.numel()
    for t in (var, grad, m, v):
        if t.numel() != n or t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
            raise ValueError("adam_step: var/grad/m/v must be contiguous CUDA float32 tensors of equal size")
    check(lib().pwc_adam_step(var.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), n, lr_t_dev.data_ptr(),
                              float(beta1), float(beta2), float(eps), float(gamma), float(grad_scale), _stream()), "pwc_adam_step")


def sumsq(x, acc, scale: float = 0.5):
    check(lib().pwc_sumsq(x.data_ptr(), x.numel(), float(scale), acc.data_ptr(), _stream()), "pwc_sumsq")
    return acc


def permute_cin(w_src, w_dst, perm):
    """w_dst[:, :, i, :] = w_src[:, :, perm[i], :] (0 where perm[i] < 0); perm: CUDA int32."""
    if w_src.dim() != 4 or w_dst.dim() != 4 or w_dst.shape[2] != perm.numel() or w_src.shape[3] != w_dst.shape[3] \
            or perm.dtype != torch.int32 or not w_src.is_contiguous() or not w_dst.is_contiguous():
        raise ValueError("permute_cin: bad arguments")
    check(lib().pwc_permute_cin(w_src.data_ptr(), w_dst.data_ptr(), perm.data_ptr(), w_dst.shape[2], w_src.shape[2],
                                w_src.shape[3], _stream()), "pwc_permute_cin")
    return w_dst


def rot_weights(w, out=None, ci_begin: int = 0, ci_count=None, ci_pad=None):
    """(3,3,Cin,Cout) -> (3,3,Cout,ci_pad): rotated by 180 degrees and transposed, restricted to input channels
    [ci_begin, ci_begin+ci_count) and zero-padded to ci_pad: the stride-1 dgrad kernel for that channel range."""
    cin, cout = w.shape[2], w.shape[3]
    ci_count = cin - ci_begin if ci_count is None else ci_count
    ci_pad = ci_count if ci_pad is None else ci_pad
    if out is None:
        out = torch.empty((3, 3, cout, ci_pad), dtype=torch.float32, device=w.device)
    elif tuple(out.shape) != (3, 3, cout, ci_pad) or not out.is_contiguous():
        raise ValueError("rot_weights: out shape mismatch")
    check(lib().pwc_conv3x3_rot_weights(w.data_ptr(), out.data_ptr(), cin, cout, ci_begin, ci_count, ci_pad, _stream()),
          "pwc_conv3x3_rot_weights")
    return out


def conv3x3_tc_f16_dgrad(dy, w_rot_packed, dx, cdx_pad: int, dilation: int = 1, mask=None, mask_alpha: float = 0.1,
                         accumulate: bool = False):
    """Stride-1 dgrad on tcgen05: dx (=|+=) conv(dy, w_rot) [* leaky'(mask)]; w_rot_packed from
    ops_tc.pack_weights_f16(rot_weights(kernel, ci_pad=cdx_pad))."""
    B, H, W, Cdx, dx_cs = _nhwc(dx, "dx")
    Bo, OH, OW, Cdy, dy_cs = _nhwc(dy, "dy")
    if (Bo, OH, OW) != (B, H, W):
        raise ValueError("conv3x3_tc_f16_dgrad: stride-1 only, dy and dx must have the same spatial shape")
    if w_rot_packed.dtype != torch.float16 or w_rot_packed.numel() * 2 != lib().pwc_conv3x3_packed_bytes_f16(Cdy, cdx_pad):
        raise ValueError("conv3x3_tc_f16_dgrad: w_rot_packed has the wrong dtype/size")
    m_cs = 0
    if mask is not None:
        Bm, Hm, Wm, Cm, m_cs = _nhwc(mask, "mask")
        if (Bm, Hm, Wm, Cm) != (B, H, W, Cdx):
            raise ValueError("conv3x3_tc_f16_dgrad: mask shape mismatch")
    check(lib().pwc_conv3x3_tc_f16_dgrad(dy.data_ptr(), dy_cs, w_rot_packed.data_ptr(), dx.data_ptr(), dx_cs, _ptr(mask), m_cs,
                                         float(mask_alpha), int(accumulate), B, H, W, Cdy, Cdx, cdx_pad, dilation, _stream()),
          "pwc_conv3x3_tc_f16_dgrad")
    return dx


def tsplit_bytes(B, H, OW, C, n_shift: int = 1) -> int:
    return int(lib().pwc_tsplit_bytes(B, H, OW, C, n_shift))


def tsplit(x, out=None, db=None, conv_input: bool = False, stride: int = 1, dilation: int = 1):
    """NHWC fp32 view -> channel-major fp16 planes [h | l*2^11], each (B, C, H, OWp): the K-major operands of the
    tensor-core wgrad.  conv_input=False: plain transpose (for dy; db += per-channel sums).  conv_input=True: three
    copies, one per horizontal tap, sampled at the conv's output columns (see pwc_tsplit_f16)."""
    B, H, W, C, x_cs = _nhwc(x, "x")
    n_shift = 3 if conv_input else 1
    OW = _same_out(W, stride) if conv_input else W
    n = tsplit_bytes(B, H, OW, C, n_shift) // 2
    if out is None:
        out = torch.empty(n, dtype=torch.float16, device=x.device)
    elif out.dtype != torch.float16 or out.numel() < n or not out.is_contiguous():
        raise ValueError("tsplit: out must be a contiguous fp16 buffer of at least tsplit_bytes()/2 elements")
    if db is not None and (conv_input or db.shape != (C,) or not db.is_contiguous()):
        raise ValueError("tsplit: db must be contiguous (C,) and is only reduced by the plain transpose")
    check(lib().pwc_tsplit_f16(x.data_ptr(), x_cs, out.data_ptr(), B, H, W, C, n_shift, stride if conv_input else 1,
                               dilation if conv_input else 1, _ptr(db), _stream()), "pwc_tsplit_f16")
    return out


def conv3x3_wgrad_tc(xT, dyT, dw, in_shape, cout: int, stride: int = 1, dilation: int = 1, cin_map=None):
    """dw += Conv2DBackpropFilter on tcgen05 from tsplit() planes; in_shape = (B, H, W, Cin) of the conv input."""
    B, H, W, Cin = in_shape
    if dw.dim() != 4 or dw.shape[:2] != (3, 3) or dw.shape[3] != cout or not dw.is_contiguous():
        raise ValueError(f"conv3x3_wgrad_tc: dw must be contiguous (3,3,Cin,{cout})")
    if cin_map is None and dw.shape[2] != Cin:
        raise ValueError("conv3x3_wgrad_tc: dw input channels differ from x (pass cin_map for concat layers)")
    check(lib().pwc_conv3x3_wgrad_tc(xT.data_ptr(), dyT.data_ptr(), dw.data_ptr(), _ptr(cin_map), dw.shape[2], B, H, W, Cin, cout,
                                     stride, dilation, _stream()), "pwc_conv3x3_wgrad_tc")
    return dw
