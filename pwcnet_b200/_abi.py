"""ctypes binding of libpwc_b200.so (C ABI declared in include/pwc_b200.h).

There is deliberately no CPU or PyTorch fallback: if the shared library is missing or a symbol
is absent, importing / calling raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpwc_b200.so")

_f32p = C.c_void_p
_i = C.c_int
_f = C.c_float
_vp = C.c_void_p
_ll = C.c_longlong

# name -> (restype, argtypes); must list every symbol include/pwc_b200.h declares
SIGNATURES = {
    "pwc_version": (_i, []),
    "pwc_last_error": (C.c_char_p, []),
    "pwc_crc32c": (C.c_uint, [_vp, _ll, C.c_uint]),
    "pwc_cost_volume_fwd": (_i, [_f32p, _i, _f32p, _i, _f32p, _i, _f32p, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "pwc_warp_cost_volume_fwd": (_i, [_f32p, _i, _f32p, _i, _f32p, _i, _f, _i, _f32p, _i, _f32p, _i,
                                      _i, _i, _i, _i, _i, _f, _vp]),
    "pwc_warp_fwd": (_i, [_f32p, _i, _f32p, _i, _f, _i, _f32p, _i, _i, _i, _i, _i, _vp]),
    "pwc_conv3x3_fwd": (_i, [_f32p, _i, _f32p, _f32p, _f32p, _i, _f32p, _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "pwc_conv3x3_tc_fwd": (_i, [_f32p, _i, _f32p, _f32p, _f32p, _i, _i, _i, _i, _i, _i, _i, _i, _f, _i, _vp]),
    "pwc_conv3x3_tc_f16_fwd": (_i, [_f32p, _i, _f32p, _f32p, _f32p, _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "pwc_conv3x3_tc_f16_head": (_i, [_f32p, _i, _f32p, _f32p, _f32p, _i, _f32p, _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "pwc_conv3x3_packed_bytes_f16": (C.c_longlong, [_i, _i]),
    "pwc_conv3x3_pack_weights_f16": (_i, [_f32p, _f32p, _i, _i, _vp]),
    "pwc_conv3x3_packed_bytes": (C.c_longlong, [_i, _i]),
    "pwc_conv3x3_pack_weights_f16_batched": (_i, [_vp, _i, _vp]),
    "pwc_conv3x3_pack_weights": (_i, [_f32p, _f32p, _i, _i, _vp]),
    "pwc_resize_bilinear_fwd": (_i, [_f32p, _i, _f32p, _i, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "pwc_lploss_level_fwd": (_i, [_f32p, _i, _i, _f32p, _i, _i, _i, _i, _f, _f, _i, _f32p, _vp]),
    "pwc_epe_fwd": (_i, [_f32p, _f32p, _i, _i, _i, _f32p, _vp]),
    "pwc_conv3x3_dgrad": (_i, [_f32p, _i, _f32p, _f32p, _i, _f32p, _i, _f, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "pwc_conv3x3_wgrad": (_i, [_f32p, _i, _f32p, _i, _f32p, _f32p, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "pwc_leaky_bwd": (_i, [_f32p, _i, _f32p, _i, _ll, _i, _f, _vp]),
    "pwc_add_strided": (_i, [_f32p, _i, _f32p, _i, _ll, _i, _f, _vp]),
    "pwc_dilate2": (_i, [_f32p, _i, _f32p, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "pwc_cost_volume_bwd": (_i, [_f32p, _i, _f32p, _i, _f32p, _i, _f32p, _i, _f32p, _i, _f32p, _i, _f32p, _i, _i,
                                 _i, _i, _i, _i, _i, _f, _vp]),
    "pwc_warp_bwd": (_i, [_f32p, _i, _f32p, _i, _f, _i, _f32p, _i, _f32p, _i, _f32p, _i, _i, _i, _i, _i, _vp]),
    "pwc_resize_bilinear_bwd": (_i, [_f32p, _i, _f32p, _i, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "pwc_lploss_level_bwd": (_i, [_f32p, _i, _i, _f32p, _i, _i, _i, _i, _f, _f, _i, _f32p, _i, _i, _vp]),
    "pwc_adam_step": (_i, [_f32p, _f32p, _f32p, _f32p, _ll, _f32p, _f, _f, _f, _f, _f, _vp]),
    "pwc_sumsq": (_i, [_f32p, _ll, _f, _f32p, _vp]),
    "pwc_permute_cin": (_i, [_f32p, _f32p, _vp, _i, _i, _i, _vp]),
    "pwc_conv3x3_rot_weights": (_i, [_f32p, _f32p, _i, _i, _i, _i, _i, _vp]),
    "pwc_tsplit_bytes": (C.c_longlong, [_i, _i, _i, _i, _i]),
    "pwc_tsplit_f16": (_i, [_f32p, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _f32p, _vp]),
    "pwc_conv3x3_wgrad_tc": (_i, [_vp, _vp, _f32p, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "pwc_conv3x3_s2d_reindex": (_i, [_f32p, _f32p, _i, _i, _vp]),
    "pwc_conv3x3_s2_tc_f16_fwd": (_i, [_f32p, _vp, _f32p, _f32p, _i, _vp, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "pwc_conv3x3_tc_f16_split_fwd": (_i, [_vp, _i, _i, _f32p, _f32p, _f32p, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "pwc_conv_first_fwd": (_i, [_vp, _i, _f32p, _f32p, _f32p, _f32p, _i, _i, _i, _i, _f, _vp]),
    "pwc_count_nonfinite": (_i, [_f32p, _i, _i, _ll, _vp, _vp]),
    "pwc_u8_to_f32_fwd": (_i, [_vp, _f32p, _ll, _f32p, _vp]),
    "pwc_split_f16_fwd": (_i, [_f32p, _i, _vp, _f32p, _i, _ll, _i, _f, _vp]),
    "pwc_warp_split_fwd": (_i, [_f32p, _i, _f32p, _i, _f, _i, _vp, _i, _i, _i, _i, _vp]),
    "pwc_cost_volume_split_fwd": (_i, [_vp, _vp, _f32p, _i, _i, _i, _i, _i, _f, _f, _vp]),
    "pwc_cost_volume_split_slot_fwd": (_i, [_vp, _vp, _f32p, _i, _i, _f32p, _i, _i, _i, _i, _f, _f, _vp]),
    "pwc_conv3x3_tc_f16_dgrad": (_i, [_f32p, _i, _f32p, _f32p, _i, _f32p, _i, _f, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
}

_lib = None


class PwcError(RuntimeError):
    pass


def lib():
    """Load the shared library once; raise loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PwcError(f"{LIB_PATH} not found: build it with `python -m pwcnet_b200.build` "
                           "(there is no CPU / PyTorch fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)   # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        if l.pwc_version() != 1:
            raise PwcError(f"ABI version mismatch: library {l.pwc_version()}, binding 1")
        _lib = l
    return _lib


LAUNCHES = 0   # C-ABI compute calls that returned success (each enqueues one kernel; bench.py's gpu_launches)


def check(rc: int, what: str) -> None:
    global LAUNCHES
    LAUNCHES += 1
    if rc != 0:
        msg = lib().pwc_last_error().decode(errors="replace")
        raise PwcError(f"{what} failed with code {rc}: {msg}")
