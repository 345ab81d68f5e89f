"""tcgen05 tensor-core convolution operators (see csrc/conv_tc.cu)."""
from __future__ import annotations

import torch

from ._abi import check, lib
from .ops import _nhwc, _stream, new_nhwc


def pack_weights(kernel_hwio: torch.Tensor, out=None) -> torch.Tensor:
    """HWIO (3,3,Cin,Cout) -> packed [2][9][Cout][Cin_pad] (tf32-exact hi plane + residual lo plane)."""
    if kernel_hwio.dim() != 4 or kernel_hwio.shape[:2] != (3, 3) or not kernel_hwio.is_cuda \
            or kernel_hwio.dtype != torch.float32 or not kernel_hwio.is_contiguous():
        raise ValueError("pack_weights: kernel must be a contiguous CUDA float32 HWIO (3,3,Cin,Cout) tensor")
    cin, cout = kernel_hwio.shape[2], kernel_hwio.shape[3]
    nbytes = lib().pwc_conv3x3_packed_bytes(cin, cout)
    if out is None:
        out = torch.empty(nbytes // 4, dtype=torch.float32, device=kernel_hwio.device)
    elif out.dtype != torch.float32 or out.numel() * 4 != nbytes or not out.is_contiguous():
        raise ValueError("pack_weights: out has the wrong dtype/size")
    check(lib().pwc_conv3x3_pack_weights(kernel_hwio.data_ptr(), out.data_ptr(), cin, cout, _stream()),
          "pwc_conv3x3_pack_weights")
    return out


def conv3x3_tc(x, w_packed, bias, cin: int, cout: int, dilation: int = 1, alpha: float = 1.0, n_split: int = 3,
               out=None, stride: int = 1):
    """3x3 SAME conv (stride 1 or 2) + bias + leaky on tcgen05 (n_split=1: TF32, 3: 3xTF32 fp32-class)."""
    B, H, W, C, x_cs = _nhwc(x, "x")
    if C != cin:
        raise ValueError(f"conv3x3_tc: x has {C} channels, weights expect {cin}")
    if bias.shape != (cout,) or not bias.is_cuda or bias.dtype != torch.float32:
        raise ValueError("conv3x3_tc: bias must be CUDA float32 (Cout,)")
    if w_packed.numel() * 4 != lib().pwc_conv3x3_packed_bytes(cin, cout):
        raise ValueError("conv3x3_tc: w_packed has the wrong size for (Cin, Cout)")
    OH, OW = -(-H // stride), -(-W // stride)
    if out is None:
        out = new_nhwc(B, OH, OW, cout, x.device)
    Bo, Ho, Wo, Co, y_cs = _nhwc(out, "out")
    if (Bo, Ho, Wo, Co) != (B, OH, OW, cout):
        raise ValueError("conv3x3_tc: out shape mismatch")
    check(lib().pwc_conv3x3_tc_fwd(x.data_ptr(), x_cs, w_packed.data_ptr(), bias.data_ptr(), out.data_ptr(), y_cs,
                                   B, H, W, cin, cout, stride, dilation, float(alpha), n_split, _stream()), "pwc_conv3x3_tc_fwd")
    return out


def pack_weights_f16(kernel_hwio: torch.Tensor, out=None) -> torch.Tensor:
    """HWIO (3,3,Cin,Cout) fp32 -> packed fp16 [2][9][Cout][Cin_pad32] (h plane + scaled-residual l plane)."""
    if kernel_hwio.dim() != 4 or kernel_hwio.shape[:2] != (3, 3) or not kernel_hwio.is_cuda \
            or kernel_hwio.dtype != torch.float32 or not kernel_hwio.is_contiguous():
        raise ValueError("pack_weights_f16: kernel must be a contiguous CUDA float32 HWIO (3,3,Cin,Cout) tensor")
    cin, cout = kernel_hwio.shape[2], kernel_hwio.shape[3]
    nbytes = lib().pwc_conv3x3_packed_bytes_f16(cin, cout)
    if out is None:
        out = torch.empty(nbytes // 2, dtype=torch.float16, device=kernel_hwio.device)
    elif out.dtype != torch.float16 or out.numel() * 2 != nbytes or not out.is_contiguous():
        raise ValueError("pack_weights_f16: out has the wrong dtype/size")
    check(lib().pwc_conv3x3_pack_weights_f16(kernel_hwio.data_ptr(), out.data_ptr(), cin, cout, _stream()),
          "pwc_conv3x3_pack_weights_f16")
    return out


class PackJobs:
    """Device job table for `pack_weights_f16_batched`: every job packs one weight tensor, all of them in one launch.
    The tensors are referenced by address: they must stay allocated (the model refreshes them in place)."""

    def __init__(self, device):
        self.device = device
        self._rows = []
        self._keep = []
        self._table = None

    def __len__(self):
        return len(self._rows)

    def _check(self, w, out, K, N):
        if not (w.is_cuda and w.dtype == torch.float32 and w.is_contiguous() and w.dim() == 4 and tuple(w.shape[:2]) == (3, 3)):
            raise ValueError("PackJobs: kernel must be a contiguous CUDA float32 HWIO (3,3,Cin,Cout) tensor")
        if out.dtype != torch.float16 or not out.is_contiguous() or out.numel() * 2 != lib().pwc_conv3x3_packed_bytes_f16(K, N):
            raise ValueError("PackJobs: out has the wrong dtype/size")

    def add_forward(self, kernel_hwio, out, cout_pad=None):
        """out = pack_weights_f16(kernel zero-padded to cout_pad output channels)."""
        K, a = kernel_hwio.shape[2], kernel_hwio.shape[3]
        N = a if cout_pad is None else cout_pad
        self._check(kernel_hwio, out, K, N)
        self._rows.append([kernel_hwio.data_ptr(), out.data_ptr(), K, N, 0, a, 0, 0])
        self._keep.append((kernel_hwio, out)); self._table = None

    def add_dgrad(self, kernel_hwio, out, ci_begin, ci_count, ci_pad):
        """out = pack_weights_f16(rot_weights(kernel, ci_begin, ci_count, ci_pad)): the stride-1 dgrad kernel."""
        a, K = kernel_hwio.shape[2], kernel_hwio.shape[3]
        if not (0 <= ci_begin and ci_count > 0 and ci_begin + ci_count <= a and ci_count <= ci_pad):
            raise ValueError("PackJobs: bad input-channel range")
        self._check(kernel_hwio, out, K, ci_pad)
        self._rows.append([kernel_hwio.data_ptr(), out.data_ptr(), K, ci_pad, 1, a, ci_begin, ci_count])
        self._keep.append((kernel_hwio, out)); self._table = None

    def run(self):
        if not self._rows:
            return
        if self._table is None:
            self._table = torch.tensor(self._rows, dtype=torch.int64, device=self.device)
        check(lib().pwc_conv3x3_pack_weights_f16_batched(self._table.data_ptr(), len(self._rows), _stream()),
              "pwc_conv3x3_pack_weights_f16_batched")


def conv3x3_tc_f16(x, w_packed, bias, cin: int, cout: int, dilation: int = 1, alpha: float = 1.0, out=None,
                   stride: int = 1):
    """3x3 SAME conv (stride 1 or 2) + bias + leaky on tcgen05 kind::f16 with the 3 x fp16 scaled-residual split."""
    B, H, W, C, x_cs = _nhwc(x, "x")
    if C != cin:
        raise ValueError(f"conv3x3_tc_f16: x has {C} channels, weights expect {cin}")
    if bias.shape != (cout,) or not bias.is_cuda or bias.dtype != torch.float32:
        raise ValueError("conv3x3_tc_f16: bias must be CUDA float32 (Cout,)")
    if w_packed.dtype != torch.float16 or w_packed.numel() * 2 != lib().pwc_conv3x3_packed_bytes_f16(cin, cout):
        raise ValueError("conv3x3_tc_f16: w_packed has the wrong dtype/size for (Cin, Cout)")
    OH, OW = -(-H // stride), -(-W // stride)
    if out is None:
        out = new_nhwc(B, OH, OW, cout, x.device)
    Bo, Ho, Wo, Co, y_cs = _nhwc(out, "out")
    if (Bo, Ho, Wo, Co) != (B, OH, OW, cout):
        raise ValueError("conv3x3_tc_f16: out shape mismatch")
    check(lib().pwc_conv3x3_tc_f16_fwd(x.data_ptr(), x_cs, w_packed.data_ptr(), bias.data_ptr(), out.data_ptr(), y_cs,
                                       B, H, W, cin, cout, stride, dilation, float(alpha), _stream()),
          "pwc_conv3x3_tc_f16_fwd")
    return out


def conv3x3_tc_f16_split(x, w_packed, bias, cin: int, cout: int, dilation: int = 1, alpha: float = 1.0, out=None, out_split=None):
    """Stride-1 conv of a conv -> conv chain with split activations (halo kernel): `x` is a float32 NHWC view or a split
    tensor (contiguous fp16 (B,H,W,2*cin): per pixel and 32-channel slice [h x32 | l x32]); the result goes to `out` (float32
    NHWC view) and/or `out_split` (contiguous fp16 (B,H,W,2*cout)).  Bit-identical to conv3x3_tc_f16 on float32 tensors."""
    if x.dtype == torch.float16:
        if x.dim() != 4 or not x.is_contiguous() or x.shape[3] != 2 * cin or cin % 32:
            raise ValueError("conv3x3_tc_f16_split: split input must be contiguous fp16 (B,H,W,2*cin), cin % 32 == 0")
        B, H, W, _ = x.shape
        x_split, x_cs = 1, 2 * cin
    else:
        B, H, W, C, x_cs = _nhwc(x, "x")
        if C != cin:
            raise ValueError(f"conv3x3_tc_f16_split: x has {C} channels, weights expect {cin}")
        x_split = 0
    if bias.shape != (cout,) or w_packed.dtype != torch.float16 or w_packed.numel() * 2 != lib().pwc_conv3x3_packed_bytes_f16(cin, cout):
        raise ValueError("conv3x3_tc_f16_split: bias / w_packed do not match (Cin, Cout)")
    if out is None and out_split is None:
        out = new_nhwc(B, H, W, cout, x.device)
    yp, y_cs = 0, 0
    if out is not None:
        Bo, Ho, Wo, Co, y_cs = _nhwc(out, "out")
        if (Bo, Ho, Wo, Co) != (B, H, W, cout):
            raise ValueError("conv3x3_tc_f16_split: out shape mismatch")
        yp = out.data_ptr()
    sp, ys_cs = 0, 0
    if out_split is not None:
        if out_split.dtype != torch.float16 or tuple(out_split.shape) != (B, H, W, 2 * cout) or not out_split.is_contiguous() or cout % 32:
            raise ValueError("conv3x3_tc_f16_split: out_split must be contiguous fp16 (B,H,W,2*cout), cout % 32 == 0")
        sp, ys_cs = out_split.data_ptr(), 2 * cout
    check(lib().pwc_conv3x3_tc_f16_split_fwd(x.data_ptr(), x_split, x_cs, w_packed.data_ptr(), bias.data_ptr(), yp, y_cs, sp, ys_cs,
                                             B, H, W, cin, cout, dilation, float(alpha), _stream()), "pwc_conv3x3_tc_f16_split_fwd")
    return out if out is not None else out_split


def s2d_reindex(k_hwio, out=None):
    """(3,3,Cin,Cout) HWIO kernel of a stride-2 conv -> the (3,3,4*Cin,Cout) kernel of the equivalent 2x2 convolution over the
    space-to-depth view (taps (0..1, 0..1); see pwc_conv3x3_s2_tc_f16_fwd).  Pack the result with pack_weights_f16."""
    if k_hwio.dim() != 4 or tuple(k_hwio.shape[:2]) != (3, 3) or k_hwio.dtype != torch.float32 or not k_hwio.is_contiguous():
        raise ValueError("s2d_reindex: kernel must be contiguous float32 (3,3,Cin,Cout)")
    cin, cout = k_hwio.shape[2], k_hwio.shape[3]
    if out is None:
        out = torch.empty((3, 3, 4 * cin, cout), dtype=torch.float32, device=k_hwio.device)
    if tuple(out.shape) != (3, 3, 4 * cin, cout) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("s2d_reindex: out must be contiguous float32 (3,3,4*Cin,Cout)")
    check(lib().pwc_conv3x3_s2d_reindex(k_hwio.data_ptr(), out.data_ptr(), cin, cout, _stream()), "pwc_conv3x3_s2d_reindex")
    return out


def conv3x3_s2_tc_f16(x, w_packed_s2d, bias, cin: int, cout: int, alpha: float = 1.0, out=None, out_split=None):
    """Stride-2 3x3 conv + bias + leaky (modules.py:62-63) on the halo kernel.  `x`: dense float32 NHWC with even H, W;
    `w_packed_s2d` = pack_weights_f16(s2d_reindex(kernel)); result (B,H/2,W/2,cout) to `out` (float32 view) and/or
    `out_split` (contiguous fp16 (B,H/2,W/2,2*cout), cout % 32 == 0).  Same numerics class as conv3x3_tc_f16(stride=2)."""
    B, H, W, C, x_cs = _nhwc(x, "x")
    if C != cin or x_cs != cin or H % 2 or W % 2 or cin % 16 or cout % 16 or cout > 128:
        raise ValueError("conv3x3_s2_tc_f16: needs a dense input, even H and W, cin % 16 == 0, cout % 16 == 0, cout <= 128")
    if bias.shape != (cout,) or w_packed_s2d.dtype != torch.float16 or \
            w_packed_s2d.numel() * 2 != lib().pwc_conv3x3_packed_bytes_f16(4 * cin, cout):
        raise ValueError("conv3x3_s2_tc_f16: bias / w_packed_s2d do not match (4*Cin, Cout)")
    OH, OW = H // 2, W // 2
    if out is None and out_split is None:
        out = new_nhwc(B, OH, OW, cout, x.device)
    yp, y_cs = 0, 0
    if out is not None:
        Bo, Ho, Wo, Co, y_cs = _nhwc(out, "out")
        if (Bo, Ho, Wo, Co) != (B, OH, OW, cout):
            raise ValueError("conv3x3_s2_tc_f16: out shape mismatch")
        yp = out.data_ptr()
    sp, ys_cs = 0, 0
    if out_split is not None:
        if out_split.dtype != torch.float16 or tuple(out_split.shape) != (B, OH, OW, 2 * cout) or not out_split.is_contiguous() or cout % 32:
            raise ValueError("conv3x3_s2_tc_f16: out_split must be contiguous fp16 (B,H/2,W/2,2*cout), cout % 32 == 0")
        sp, ys_cs = out_split.data_ptr(), 2 * cout
    check(lib().pwc_conv3x3_s2_tc_f16_fwd(x.data_ptr(), w_packed_s2d.data_ptr(), bias.data_ptr(), yp, y_cs, sp, ys_cs,
                                          B, H, W, cin, cout, float(alpha), _stream()), "pwc_conv3x3_s2_tc_f16_fwd")
    return out if out is not None else out_split


def conv3x3_tc_f16_head(x, w_packed_pad, bias_pad, cin: int, cout: int, cout_pad: int, dilation: int = 1, alpha: float = 1.0,
                        residual=None, out=None):
    """Narrow-output conv (the 2-channel flow heads, modules.py:274-277, 325-326) on tcgen05: kernel and bias are
    zero-padded to `cout_pad` (multiple of 16) output channels, only `cout` are stored; `residual` is added after the
    activation."""
    B, H, W, C, x_cs = _nhwc(x, "x")
    if C != cin or cout_pad % 16 or cout > cout_pad:
        raise ValueError("conv3x3_tc_f16_head: channel mismatch")
    if bias_pad.shape != (cout_pad,) or w_packed_pad.numel() * 2 != lib().pwc_conv3x3_packed_bytes_f16(cin, cout_pad):
        raise ValueError("conv3x3_tc_f16_head: bias / packed kernel must be padded to cout_pad")
    if out is None:
        out = new_nhwc(B, H, W, cout, x.device)
    Bo, Ho, Wo, Co, y_cs = _nhwc(out, "out")
    if (Bo, Ho, Wo, Co) != (B, H, W, cout):
        raise ValueError("conv3x3_tc_f16_head: out shape mismatch")
    rp, r_cs = None, 0
    if residual is not None:
        Br, Hr, Wr, Cr, r_cs = _nhwc(residual, "residual")
        if (Br, Hr, Wr, Cr) != (B, H, W, cout):
            raise ValueError("conv3x3_tc_f16_head: residual shape mismatch")
        rp = residual.data_ptr()
    check(lib().pwc_conv3x3_tc_f16_head(x.data_ptr(), x_cs, w_packed_pad.data_ptr(), bias_pad.data_ptr(), rp, r_cs, out.data_ptr(),
                                        y_cs, B, H, W, cin, cout, cout_pad, dilation, float(alpha), _stream()), "pwc_conv3x3_tc_f16_head")
    return out
