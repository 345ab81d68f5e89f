"""Functional operators over torch CUDA tensors (NHWC float32) that launch the sm_100a kernels of
libpwc_b200.so through the C ABI.  PyTorch is the container (device memory, streams) only.

Every operator accepts "pixel-strided" NHWC views: the channel dim is contiguous and the spatial /
batch strides are those of a dense NHW(Cs) buffer with Cs >= C, i.e. `buf[..., a:b]` slots of a
wider concat buffer are valid inputs and outputs (this replaces tf.concat, modules.py:262-264,305).
Arguments are validated in Python before any launch; failures raise (no fallbacks).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ._abi import PwcError, check, lib

WARP_TYPES = {"bilinear": 0, "nearest": 1}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _nhwc(t: torch.Tensor, name: str) -> Tuple[int, int, int, int, int]:
    """Validate a pixel-strided NHWC float32 CUDA view -> (B, H, W, C, channel_stride)."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t)}")
    if not t.is_cuda:
        raise PwcError(f"{name}: tensor must live on a CUDA device (there is no CPU path)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: dtype must be float32, got {t.dtype}")
    if t.dim() != 4:
        raise ValueError(f"{name}: expected NHWC rank-4 tensor, got shape {tuple(t.shape)}")
    B, H, W, C = t.shape
    if min(B, H, W, C) <= 0:
        raise ValueError(f"{name}: empty tensor {tuple(t.shape)}")
    sb, sh, sw, sc = t.stride()
    cs = sw if W > 1 else (sh if H > 1 else (sb if B > 1 else C))
    ok = (sc == 1 or C == 1) and cs >= C
    ok = ok and (W == 1 or sw == cs) and (H == 1 or sh == W * cs) and (B == 1 or sb == H * W * cs)
    if not ok:
        raise ValueError(f"{name}: not a pixel-strided NHWC view (shape {tuple(t.shape)}, strides {t.stride()})")
    return B, H, W, C, cs


def _same_out(n: int, stride: int) -> int:
    return -(-n // stride)


def new_nhwc(B, H, W, C, device, cs: Optional[int] = None) -> torch.Tensor:
    cs = C if cs is None else cs
    return torch.empty((B, H, W, cs), dtype=torch.float32, device=device)[..., :C]


def cost_volume(f0, f1, search_range: int = 4, alpha: float = 0.1, out=None, f0_copy=None):
    """CostVolumeLayer.__call__ (modules.py:189-204)."""
    B, H, W, C, f0_cs = _nhwc(f0, "features_0")
    B1, H1, W1, C1, f1_cs = _nhwc(f1, "features_0from1")
    if (B, H, W, C) != (B1, H1, W1, C1):
        raise ValueError(f"cost_volume: shape mismatch {tuple(f0.shape)} vs {tuple(f1.shape)}")
    nd = (2 * search_range + 1) ** 2
    if out is None:
        out = new_nhwc(B, H, W, nd, f0.device)
    Bo, Ho, Wo, Co, out_cs = _nhwc(out, "out")
    if (Bo, Ho, Wo, Co) != (B, H, W, nd):
        raise ValueError(f"cost_volume: out has shape {tuple(out.shape)}, expected {(B, H, W, nd)}")
    cp, cp_cs = None, 0
    if f0_copy is not None:
        Bc, Hc, Wc, Cc, cp_cs = _nhwc(f0_copy, "f0_copy")
        if (Bc, Hc, Wc, Cc) != (B, H, W, C):
            raise ValueError("cost_volume: f0_copy shape mismatch")
        cp = f0_copy.data_ptr()
    check(lib().pwc_cost_volume_fwd(f0.data_ptr(), f0_cs, f1.data_ptr(), f1_cs, out.data_ptr(), out_cs, cp, cp_cs,
                                    B, H, W, C, search_range, alpha, _stream()), "pwc_cost_volume_fwd")
    return out


def warp_cost_volume(f0, f1, flow, flow_scale: float = 1.0, warp_type: str = "bilinear", search_range: int = 4,
                     alpha: float = 0.1, out=None, f0_copy=None):
    """WarpingLayer + CostVolumeLayer fused (model.py:109-112)."""
    if warp_type not in WARP_TYPES:
        raise AssertionError(f"warp_type must be one of {list(WARP_TYPES)}")   # modules.py:149
    B, H, W, C, f0_cs = _nhwc(f0, "features_0")
    B1, H1, W1, C1, f1_cs = _nhwc(f1, "features_1")
    Bf, Hf, Wf, Cf, fl_cs = _nhwc(flow, "flow")
    if (B, H, W, C) != (B1, H1, W1, C1) or (Bf, Hf, Wf, Cf) != (B, H, W, 2):
        raise ValueError("warp_cost_volume: shape mismatch")
    nd = (2 * search_range + 1) ** 2
    if out is None:
        out = new_nhwc(B, H, W, nd, f0.device)
    Bo, Ho, Wo, Co, out_cs = _nhwc(out, "out")
    if (Bo, Ho, Wo, Co) != (B, H, W, nd):
        raise ValueError("warp_cost_volume: out shape mismatch")
    cp, cp_cs = None, 0
    if f0_copy is not None:
        Bc, Hc, Wc, Cc, cp_cs = _nhwc(f0_copy, "f0_copy")
        if (Bc, Hc, Wc, Cc) != (B, H, W, C):
            raise ValueError("warp_cost_volume: f0_copy shape mismatch")
        cp = f0_copy.data_ptr()
    check(lib().pwc_warp_cost_volume_fwd(f0.data_ptr(), f0_cs, f1.data_ptr(), f1_cs, flow.data_ptr(), fl_cs,
                                         float(flow_scale), WARP_TYPES[warp_type], out.data_ptr(), out_cs, cp, cp_cs,
                                         B, H, W, C, search_range, alpha, _stream()), "pwc_warp_cost_volume_fwd")
    return out


def warp(x, flow, flow_scale: float = 1.0, warp_type: str = "bilinear", out=None):
    """WarpingLayer.__call__ (modules.py:144-154)."""
    assert warp_type in ["nearest", "bilinear"]   # modules.py:149
    B, H, W, C, x_cs = _nhwc(x, "x")
    Bf, Hf, Wf, Cf, fl_cs = _nhwc(flow, "flow")
    if (Bf, Hf, Wf, Cf) != (B, H, W, 2):
        raise ValueError(f"warp: flow shape {tuple(flow.shape)} does not match x {tuple(x.shape)}")
    if out is None:
        out = new_nhwc(B, H, W, C, x.device)
    Bo, Ho, Wo, Co, out_cs = _nhwc(out, "out")
    if (Bo, Ho, Wo, Co) != (B, H, W, C):
        raise ValueError("warp: out shape mismatch")
    check(lib().pwc_warp_fwd(x.data_ptr(), x_cs, flow.data_ptr(), fl_cs, float(flow_scale), WARP_TYPES[warp_type],
                             out.data_ptr(), out_cs, B, H, W, C, _stream()), "pwc_warp_fwd")
    return out


def conv3x3(x, kernel, bias, stride: int = 1, dilation: int = 1, alpha: float = 1.0, residual=None, out=None):
    """tf.layers.Conv2D(filters,(3,3),strides,'same',dilation_rate) [+ leaky_relu(alpha)] [+ residual].
    kernel is HWIO (3,3,Cin,Cout) as stored in the reference checkpoints; alpha=1.0 = no activation."""
    B, H, W, Cin, x_cs = _nhwc(x, "x")
    if kernel.shape[:3] != (3, 3, Cin) or kernel.dim() != 4 or not kernel.is_contiguous() or not kernel.is_cuda \
            or kernel.dtype != torch.float32:
        raise ValueError(f"conv3x3: kernel must be contiguous CUDA float32 HWIO (3,3,{Cin},Cout), got {tuple(kernel.shape)}")
    Cout = kernel.shape[3]
    if bias.shape != (Cout,) or not bias.is_cuda or bias.dtype != torch.float32 or not bias.is_contiguous():
        raise ValueError("conv3x3: bias must be contiguous CUDA float32 of shape (Cout,)")
    OH, OW = _same_out(H, stride), _same_out(W, stride)
    if out is None:
        out = new_nhwc(B, OH, OW, Cout, x.device)
    Bo, Ho, Wo, Co, y_cs = _nhwc(out, "out")
    if (Bo, Ho, Wo, Co) != (B, OH, OW, Cout):
        raise ValueError(f"conv3x3: out has shape {tuple(out.shape)}, expected {(B, OH, OW, Cout)}")
    rp, r_cs = None, 0
    if residual is not None:
        Br, Hr, Wr, Cr, r_cs = _nhwc(residual, "residual")
        if (Br, Hr, Wr, Cr) != (B, OH, OW, Cout):
            raise ValueError("conv3x3: residual shape mismatch")
        rp = residual.data_ptr()
    check(lib().pwc_conv3x3_fwd(x.data_ptr(), x_cs, kernel.data_ptr(), bias.data_ptr(), rp, r_cs, out.data_ptr(), y_cs,
                                B, H, W, Cin, Cout, stride, dilation, float(alpha), _stream()), "pwc_conv3x3_fwd")
    return out


def resize_bilinear(x, out_h: int, out_w: int, mul: float = 1.0, out=None):
    """tf.image.resize_bilinear(x, (out_h,out_w)) (TF-1.8 legacy, align_corners=False) * mul."""
    B, H, W, C, x_cs = _nhwc(x, "x")
    if out is None:
        out = new_nhwc(B, out_h, out_w, C, x.device)
    Bo, Ho, Wo, Co, y_cs = _nhwc(out, "out")
    if (Bo, Ho, Wo, Co) != (B, out_h, out_w, C):
        raise ValueError("resize_bilinear: out shape mismatch")
    check(lib().pwc_resize_bilinear_fwd(x.data_ptr(), x_cs, out.data_ptr(), y_cs, B, H, W, C, out_h, out_w,
                                        float(mul), _stream()), "pwc_resize_bilinear_fwd")
    return out


def lploss_level(flows_gt, fs, weight: float, acc, gt_div: float = 20.0, ord: int = 2):
    """acc += weight * L{ord}loss(resize_nearest(flows_gt/gt_div, fs.shape), fs)   (losses.py:20-29)."""
    B, H, W, C, g_cs = _nhwc(flows_gt, "flows_gt")
    Bf, h, w, Cf, f_cs = _nhwc(fs, "flows")
    if C != 2 or Cf != 2 or B != Bf or g_cs != 2:
        raise ValueError("lploss_level: flows_gt must be dense (B,H,W,2) and flows (B,h,w,2)")
    check(lib().pwc_lploss_level_fwd(flows_gt.data_ptr(), H, W, fs.data_ptr(), f_cs, h, w, B, float(gt_div),
                                     float(weight), int(ord), acc.data_ptr(), _stream()), "pwc_lploss_level_fwd")
    return acc


def epe(flows_gt, flows, acc):
    """acc += EPE(flows_gt, flows)   (losses.py:11-13)."""
    B, H, W, C, g_cs = _nhwc(flows_gt, "flows_gt")
    B2, H2, W2, C2, f_cs = _nhwc(flows, "flows")
    if (B, H, W, C) != (B2, H2, W2, C2) or C != 2 or g_cs != 2 or f_cs != 2:
        raise ValueError("epe: flows_gt and flows must be dense tensors of the same (B,H,W,2) shape")
    check(lib().pwc_epe_fwd(flows_gt.data_ptr(), flows.data_ptr(), B, H, W, acc.data_ptr(), _stream()), "pwc_epe_fwd")
    return acc


# ------------------------------------------------------------------ split-fp16 fast path of the cost volume
def split_f16(x, out=None, copy=None, scale: float = 1.0):
    """fp32 NHWC view -> split tensor (B,H,W,2C) fp16: per pixel and 32-channel slice [h | l], h = fp16(x), l = fp16(x-h).
    `copy` (fp32 view of the same shape, e.g. the f0 slot of a concat buffer) also receives x."""
    B, H, W, C, x_cs = _nhwc(x, "x")
    if C % 32:
        raise ValueError("split_f16: C must be a multiple of 32")
    if out is None:
        out = torch.empty((B, H, W, 2 * C), dtype=torch.float16, device=x.device)
    elif out.dtype != torch.float16 or tuple(out.shape) != (B, H, W, 2 * C) or not out.is_contiguous():
        raise ValueError("split_f16: out must be a contiguous fp16 (B,H,W,2C) tensor")
    cp, cp_cs = None, 0
    if copy is not None:
        Bc, Hc, Wc, Cc, cp_cs = _nhwc(copy, "copy")
        if (Bc, Hc, Wc, Cc) != (B, H, W, C):
            raise ValueError("split_f16: copy shape mismatch")
        cp = copy.data_ptr()
    check(lib().pwc_split_f16_fwd(x.data_ptr(), x_cs, out.data_ptr(), cp, cp_cs, B * H * W, C, float(scale), _stream()),
          "pwc_split_f16_fwd")
    return out


def warp_split(x, flow, flow_scale: float = 1.0, warp_type: str = "bilinear", out=None):
    """WarpingLayer.__call__ (modules.py:144-154) with the result written as a split tensor."""
    assert warp_type in ["nearest", "bilinear"]   # modules.py:149
    B, H, W, C, x_cs = _nhwc(x, "x")
    Bf, Hf, Wf, Cf, fl_cs = _nhwc(flow, "flow")
    if (Bf, Hf, Wf, Cf) != (B, H, W, 2) or C % 32:
        raise ValueError("warp_split: flow shape mismatch or C not a multiple of 32")
    if out is None:
        out = torch.empty((B, H, W, 2 * C), dtype=torch.float16, device=x.device)
    elif out.dtype != torch.float16 or tuple(out.shape) != (B, H, W, 2 * C) or not out.is_contiguous():
        raise ValueError("warp_split: out must be a contiguous fp16 (B,H,W,2C) tensor")
    check(lib().pwc_warp_split_fwd(x.data_ptr(), x_cs, flow.data_ptr(), fl_cs, float(flow_scale), WARP_TYPES[warp_type],
                                   out.data_ptr(), B, H, W, C, _stream()), "pwc_warp_split_fwd")
    return out


def cost_volume_split(f0s, f1s, alpha: float = 0.1, out=None, prescaled: bool = False, slot=False, tail=None):
    """CostVolumeLayer.__call__ (modules.py:189-204, search_range 4) from split operands, on the tensor cores.
    prescaled=True: f0s was produced by split_f16(..., scale=1/C), the kernel skips the 1/C multiply.
    slot=True: `out` is the head of a concat-buffer pixel row and the kernel writes whole 32-byte sectors: words [0,81) =
    cost volume, [81,83) = `tail` (dense (B,H,W,2), e.g. the up-sampled flow; zeros if None), [83,88) = zeros; `out` must be
    the (B,H,W,81) view at channel 0 of a 32-byte aligned buffer whose pixel pitch is a multiple of 8 floats, >= 88."""
    slot = 88 if slot else 0
    for t, nm in ((f0s, "f0s"), (f1s, "f1s")):
        if t.dtype != torch.float16 or t.dim() != 4 or not t.is_cuda or not t.is_contiguous():
            raise ValueError(f"cost_volume_split: {nm} must be a contiguous CUDA fp16 (B,H,W,2C) split tensor")
    if f0s.shape != f1s.shape or f0s.shape[3] % 64:
        raise ValueError("cost_volume_split: operand shapes differ or 2C is not a multiple of 64")
    B, H, W, C2 = f0s.shape
    if out is None:
        out = new_nhwc(B, H, W, 81, f0s.device, cs=slot if slot else None)
    Bo, Ho, Wo, Co, out_cs = _nhwc(out, "out")
    if (Bo, Ho, Wo, Co) != (B, H, W, 81):
        raise ValueError("cost_volume_split: out shape mismatch")
    scale = 1.0 if prescaled else 2.0 / C2
    if slot:
        tp = 0
        if tail is not None:
            if tuple(tail.shape) != (B, H, W, 2) or tail.dtype != torch.float32 or not tail.is_cuda or not tail.is_contiguous():
                raise ValueError("cost_volume_split: tail must be a contiguous CUDA float32 (B,H,W,2) tensor")
            tp = tail.data_ptr()
        check(lib().pwc_cost_volume_split_slot_fwd(f0s.data_ptr(), f1s.data_ptr(), out.data_ptr(), out_cs, slot, tp, B, H, W, C2 // 2,
                                                   scale, float(alpha), _stream()), "pwc_cost_volume_split_slot_fwd")
    else:
        if tail is not None:
            raise ValueError("cost_volume_split: tail needs slot=True")
        check(lib().pwc_cost_volume_split_fwd(f0s.data_ptr(), f1s.data_ptr(), out.data_ptr(), out_cs, B, H, W, C2 // 2,
                                              scale, float(alpha), _stream()), "pwc_cost_volume_split_fwd")
    return out


_U8_LUT = {}


def u8_lut(device) -> torch.Tensor:
    """float32(float64(v) / 255.0) for v = 0..255: the values the reference's `images/255.0` feed holds."""
    key = str(device)
    if key not in _U8_LUT:
        import numpy as np
        _U8_LUT[key] = torch.from_numpy((np.arange(256, dtype=np.float64) / 255.0).astype(np.float32)).to(device)
    return _U8_LUT[key]


def u8_to_f32(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """uint8 images -> float32 in [0,1], bit-identical to the reference's host-side /255.0 (test.py:31-33)."""
    if not (x.is_cuda and out.is_cuda) or x.dtype != torch.uint8 or out.dtype != torch.float32:
        raise TypeError("u8_to_f32: x must be a CUDA uint8 tensor and out a CUDA float32 tensor")
    if not (x.is_contiguous() and out.is_contiguous()) or x.numel() != out.numel():
        raise ValueError("u8_to_f32: x and out must be contiguous and of equal size")
    check(lib().pwc_u8_to_f32_fwd(x.data_ptr(), out.data_ptr(), x.numel(), u8_lut(x.device).data_ptr(), _stream()),
          "pwc_u8_to_f32_fwd")
    return out


def count_nonfinite(x: torch.Tensor, count: torch.Tensor) -> torch.Tensor:
    """count (int32 CUDA scalar) += number of non-finite values of the pixel-strided NHWC tensor x."""
    B, H, W, C, cs = _nhwc(x, "x")
    if count.dtype != torch.int32 or not count.is_cuda or count.numel() != 1:
        raise TypeError("count_nonfinite: count must be a CUDA int32 scalar")
    check(lib().pwc_count_nonfinite(x.data_ptr(), cs, C, B * H * W, count.data_ptr(), _stream()), "pwc_count_nonfinite")
    return count


def conv_first(x: torch.Tensor, kernel: torch.Tensor, bias: torch.Tensor, alpha: float = 0.1, out=None) -> torch.Tensor:
    """First pyramid conv (3 -> 16, stride 2, SAME, + bias + leaky; modules.py:62-63) from dense float32 RGB/255 or uint8
    RGB images (B,H,W,3); uint8 values pass through the reference's /255.0 table on the fly."""
    if not x.is_cuda or x.dim() != 4 or x.shape[3] != 3 or not x.is_contiguous() or x.dtype not in (torch.float32, torch.uint8):
        raise ValueError("conv_first: x must be a contiguous CUDA (B,H,W,3) float32 or uint8 tensor")
    if tuple(kernel.shape) != (3, 3, 3, 16) or tuple(bias.shape) != (16,) or not kernel.is_contiguous():
        raise ValueError("conv_first: kernel must be (3,3,3,16) HWIO and bias (16,)")
    B, H, W, _ = x.shape
    if H % 2 or W % 4:
        raise ValueError("conv_first: H must be even and W a multiple of 4")
    if out is None:
        out = new_nhwc(B, H // 2, W // 2, 16, x.device)
    Bo, Ho, Wo, Co, y_cs = _nhwc(out, "out")
    if (Bo, Ho, Wo, Co) != (B, H // 2, W // 2, 16):
        raise ValueError("conv_first: out shape mismatch")
    u8 = x.dtype == torch.uint8
    check(lib().pwc_conv_first_fwd(x.data_ptr(), int(u8), u8_lut(x.device).data_ptr() if u8 else 0, kernel.data_ptr(), bias.data_ptr(),
                                   out.data_ptr(), y_cs, B, H, W, float(alpha), _stream()), "pwc_conv_first_fwd")
    return out
