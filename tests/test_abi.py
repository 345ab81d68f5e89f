"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/pwc_b200.h declares; the host layer validates arguments and fails loudly without
a GPU (no silent fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "pwc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pwc_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported(lib_built):
    lib = ctypes.CDLL(lib_built)
    syms = _declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/pwc_b200.h but not exported"


def test_binding_covers_header(lib_built):
    from pwcnet_b200 import _abi
    assert sorted(_abi.SIGNATURES) == _declared_symbols()
    l = _abi.lib()
    assert l.pwc_version() == 1


def test_argument_errors_do_not_launch(lib_built):
    """Bad arguments are rejected by the C entry points with PWC_E_* codes (no CUDA call made)."""
    from pwcnet_b200 import _abi
    l = _abi.lib()
    assert l.pwc_cost_volume_fwd(None, 32, None, 32, None, 81, None, 0, 1, 8, 8, 32, 4, 0.1, None) == -1
    assert b"null" in l.pwc_last_error()
    # misaligned channel count
    assert l.pwc_cost_volume_fwd(16, 30, 16, 30, 16, 81, None, 0, 1, 8, 8, 30, 4, 0.1, None) == -2
    assert l.pwc_conv3x3_fwd(16, 3, 16, 16, None, 0, 16, 16, 1, 8, 8, 3, 16, 0, 1, 0.1, None) == -1   # stride 0
    assert l.pwc_warp_fwd(16, 32, 16, 2, 1.0, 7, 16, 32, 1, 8, 8, 32, None) == -1                       # warp_type 7
    with pytest.raises(_abi.PwcError):
        _abi.check(-1, "x")


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib_built):
    import pwcnet_b200 as P
    with pytest.raises(P.ops.PwcError):
        P.PWCDCNet()
    with pytest.raises(P.ops.PwcError):
        P.ops.cost_volume(torch.zeros(1, 8, 8, 32), torch.zeros(1, 8, 8, 32))


def test_layer_table_matches_reference_variable_contract():
    import pwcnet_b200 as P
    rows = P.layer_table()
    assert len(rows) == 55
    assert sum(9 * a * b + b for _, a, b in rows) == 5029868
    assert rows[0] == ("pwcdcnet/fp_extractor/conv2d", 3, 16)
    assert rows[17] == ("pwcdcnet/fp_extractor/conv2d_17", 192, 192)
    assert rows[18] == ("pwcdcnet/optflow_0/conv2d", 273, 128)
    assert rows[-1] == ("pwcdcnet/context/conv2d_6", 32, 2)
    assert sum(9 * a * b + b for _, a, b in P.layer_table(use_dc=True)) == 40182338
    W = P.glorot_init(0)
    assert set(W) == {r[0] + s for r in rows for s in ("/kernel", "/bias")}


def test_checkpoint_reader_roundtrip(tmp_path):
    """Write a tiny SSTable index + data file in the TF bundle format and read it back."""
    import struct
    from pwcnet_b200 import checkpoint as ck

    def varint(v):
        out = b""
        while True:
            b = v & 0x7F; v >>= 7
            out += bytes([b | (0x80 if v else 0)])
            if not v:
                return out

    def entry(dtype, shape, off, size):
        shp = b"".join(b"\x12" + varint(len(d)) + d for d in [b"\x08" + varint(s) for s in shape])
        return b"\x08" + varint(dtype) + b"\x12" + varint(len(shp)) + shp + b"\x20" + varint(off) + b"\x28" + varint(size)

    a = np.arange(24, dtype=np.float32).reshape(3, 2, 4)
    b = np.array([7], np.int32)
    data = a.tobytes() + b.tobytes()
    kv = [(b"", b"\x08\x01"), (b"Variable", entry(3, (1,), 96, 4)), (b"pwcdcnet/x/kernel", entry(1, (3, 2, 4), 0, 96))]
    block, prev = b"", b""
    for k, v in kv:
        shared = 0
        while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
            shared += 1
        block += varint(shared) + varint(len(k) - shared) + varint(len(v)) + k[shared:] + v
        prev = k
    block += struct.pack("<II", 0, 1)
    data_block = block + b"\x00" + b"\x00" * 4
    handle = varint(0) + varint(len(block))
    iblock = varint(0) + varint(1) + varint(len(handle)) + b"z" + handle + struct.pack("<II", 0, 1)
    index_block = iblock + b"\x00" + b"\x00" * 4
    meta = struct.pack("<I", 0) + struct.pack("<I", 1)
    meta_block = meta + b"\x00" + b"\x00" * 4
    off_meta = len(data_block)
    off_index = off_meta + len(meta_block)
    footer = varint(off_meta) + varint(len(meta)) + varint(off_index) + varint(len(iblock))
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57)
    prefix = str(tmp_path / "m.ckpt")
    open(prefix + ".index", "wb").write(data_block + meta_block + index_block + footer)
    open(prefix + ".data-00000-of-00001", "wb").write(data)
    W = ck.load_checkpoint(prefix)
    assert list(W) == ["pwcdcnet/x/kernel"]
    np.testing.assert_array_equal(W["pwcdcnet/x/kernel"], a)
    assert ck.read_scalar(prefix, "Variable")[0] == 7
