"""Reader for TensorFlow-1.x checkpoint bundles (`*.ckpt.index` + `*.ckpt.data-00000-of-00001`)
without TensorFlow, so that checkpoints written by the reference's `tf.train.Saver`
(train.py:95,166; restored at train.py:97-99 and test.py:40-42) load straight into
`PWCDCNet.load_weights`.  Format notes: SURVEY.md 9.8.

The `.index` file is a LevelDB-style SSTable (uncompressed blocks, prefix-compressed keys);
key "" holds the BundleHeaderProto, every other key is a variable name whose value is a
BundleEntryProto {1:dtype 2:shape 3:shard_id 4:offset 5:size 6:crc32c}.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Tuple

import numpy as np

_MAGIC = 0xDB4775248B80FB57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}


def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _block_entries(block: bytes):
    """Yield (key, value) from one SSTable block (restart array stripped)."""
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 * (n_restarts + 1)
    pos, key = 0, b""
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def _read_block(buf: bytes, off: int, size: int) -> bytes:
    if buf[off + size] != 0:
        raise ValueError("compressed SSTable blocks are not supported")
    return buf[off:off + size]


def _parse_proto(buf: bytes) -> Dict[int, list]:
    """Minimal protobuf wire parser: field number -> list of raw values."""
    out: Dict[int, list] = {}
    pos = 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]; pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]; pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]; pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        out.setdefault(field, []).append(v)
    return out


def _entry(value: bytes):
    p = _parse_proto(value)
    dtype = p.get(1, [0])[0]
    shape = []
    if 2 in p:
        sp = _parse_proto(p[2][0])
        for dim in sp.get(2, []):
            d = _parse_proto(dim)
            shape.append(d.get(1, [0])[0])
    return dtype, tuple(shape), p.get(3, [0])[0], p.get(4, [0])[0], p.get(5, [0])[0]


def list_variables(prefix: str) -> Dict[str, Tuple[np.dtype, Tuple[int, ...], int, int]]:
    """name -> (dtype, shape, offset, size) for every tensor in the bundle at `prefix`."""
    with open(prefix + ".index", "rb") as f:
        buf = f.read()
    footer = buf[-48:]
    if struct.unpack_from("<Q", footer, 40)[0] != _MAGIC:
        raise ValueError(f"{prefix}.index is not a TF checkpoint index (bad magic)")
    pos = 0
    _, pos = _varint(footer, pos); _, pos = _varint(footer, pos)       # metaindex handle
    ioff, pos = _varint(footer, pos); isz, pos = _varint(footer, pos)  # index handle
    out = {}
    for _, handle in _block_entries(_read_block(buf, ioff, isz)):
        boff, p = _varint(handle, 0)
        bsz, p = _varint(handle, p)
        for key, value in _block_entries(_read_block(buf, boff, bsz)):
            if key == b"":
                continue
            dtype, shape, shard, off, size = _entry(value)
            if shard != 0:
                raise ValueError("multi-shard bundles are not supported")
            if dtype not in _DTYPES:
                continue
            out[key.decode()] = (np.dtype(_DTYPES[dtype]), shape, off, size)
    return out


def load_checkpoint(prefix: str, name_filter: str = "pwcdcnet", include_slots: bool = False) -> Dict[str, np.ndarray]:
    """Load tensors whose name contains `name_filter` (the reference's `model.vars`
    filter, model.py:136-138).  Adam slot variables (`.../Adam`, `.../Adam_1`) are skipped
    unless `include_slots`."""
    vars_ = list_variables(prefix)
    data_path = prefix + ".data-00000-of-00001"
    if not os.path.exists(data_path):
        raise FileNotFoundError(data_path)
    out = {}
    with open(data_path, "rb") as f:
        for name, (dtype, shape, off, size) in sorted(vars_.items()):
            if name_filter not in name:
                continue
            if not include_slots and (name.endswith("/Adam") or name.endswith("/Adam_1")):
                continue
            f.seek(off)
            arr = np.frombuffer(f.read(size), dtype=dtype)
            out[name] = arr.reshape(shape).copy()
    return out


def read_scalar(prefix: str, name: str):
    """e.g. read_scalar(prefix, 'Variable') -> global_step."""
    dtype, shape, off, size = list_variables(prefix)[name]
    with open(prefix + ".data-00000-of-00001", "rb") as f:
        f.seek(off)
        return np.frombuffer(f.read(size), dtype=dtype).reshape(shape)


# ------------------------------------------------------------------------------------------------ writer
# tf.train.Saver.save (train.py:95,166) writes `<prefix>.index` + `<prefix>.data-00000-of-00001` (+ a .meta GraphDef the
# reference's restore paths never read: train.py:97-99 / test.py:40-42 call saver.restore on a graph they built
# themselves).  save_checkpoint() produces the same two files byte for byte (verified against the reference's own
# checkpoints in tests/test_checkpoint_writer.py): tensors in byte-wise name order, BundleEntryProto per tensor with the
# masked CRC32C TensorFlow verifies on restore, one LevelDB-format table (restart interval 16, 256 KiB blocks, no
# compression, masked CRC32C block trailers).
_DTYPE_ENUM = {np.dtype(np.float32): 1, np.dtype(np.float64): 2, np.dtype(np.int32): 3, np.dtype(np.int64): 9}
_CRC_TABLE = None


def _crc32c_py(data: bytes, crc: int = 0) -> int:
    """CRC-32C (Castagnoli), table-driven; only used when libpwc_b200.so is not built."""
    global _CRC_TABLE
    if _CRC_TABLE is None:
        tab = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            tab.append(c)
        _CRC_TABLE = tab
    tab = _CRC_TABLE
    c = crc ^ 0xFFFFFFFF
    for b in data:
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def crc32c(data) -> int:
    """CRC-32C of a bytes-like object: `pwc_crc32c` of the C-ABI library (slice-by-8, host code) when it is built."""
    mv = memoryview(data).cast("B")
    if len(mv) >= 4096:
        try:
            from . import _abi
            import ctypes
            buf = np.frombuffer(mv, dtype=np.uint8)
            return int(_abi.lib().pwc_crc32c(ctypes.c_void_p(buf.ctypes.data), len(mv), 0)) & 0xFFFFFFFF
        except Exception:
            pass
    return _crc32c_py(bytes(mv))


def _mask_crc(c: int) -> int:
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def _enc_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _entry_proto(arr: np.ndarray, offset: int) -> bytes:
    dims = b"".join(b"\x12" + _enc_varint(len(d)) + d for d in (b"\x08" + _enc_varint(int(s)) if s else b"" for s in arr.shape))
    out = b"\x08" + _enc_varint(_DTYPE_ENUM[arr.dtype]) + b"\x12" + _enc_varint(len(dims)) + dims
    if offset:
        out += b"\x20" + _enc_varint(offset)
    out += b"\x28" + _enc_varint(arr.nbytes)
    out += b"\x35" + struct.pack("<I", _mask_crc(crc32c(arr.tobytes() if arr.nbytes < 4096 else arr)))
    return out


def _build_block(kvs) -> bytes:
    """One table block: prefix-compressed entries, restart point every 16 entries, restart array + count."""
    out, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(kvs):
        shared = 0
        if i % 16 == 0:
            restarts.append(len(out))
        else:
            n = min(len(prev), len(k))
            while shared < n and prev[shared] == k[shared]:
                shared += 1
        out += _enc_varint(shared) + _enc_varint(len(k) - shared) + _enc_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    out += struct.pack(f"<{len(restarts)}I", *restarts) + struct.pack("<I", len(restarts))
    return bytes(out)


def _with_trailer(block: bytes) -> bytes:
    return block + b"\x00" + struct.pack("<I", _mask_crc(crc32c(block + b"\x00")))


def _short_successor(key: bytes) -> bytes:
    for i, b in enumerate(key):
        if b != 0xFF:
            return key[:i] + bytes([b + 1])
    return key


def _shortest_separator(a: bytes, b: bytes) -> bytes:
    n = min(len(a), len(b))
    i = 0
    while i < n and a[i] == b[i]:
        i += 1
    if i < n and a[i] < 0xFF and a[i] + 1 < b[i]:
        return a[:i] + bytes([a[i] + 1])
    return a


def save_checkpoint(prefix: str, tensors: Dict[str, np.ndarray], block_size: int = 262144) -> None:
    """Write `tensors` (name -> array; float32 / float64 / int32 / int64) as a TensorFlow-1.x checkpoint bundle that
    tf.train.Saver.restore -- and load_checkpoint() above -- read."""
    names = sorted(tensors, key=lambda s: s.encode())
    kvs = [(b"", b"\x08\x01\x1a\x02\x08\x01")]       # BundleHeaderProto{num_shards: 1, version{producer: 1}}
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for name in names:
            arr = np.asarray(tensors[name], order="C")        # (ascontiguousarray would turn scalars into shape (1,))
            if arr.dtype not in _DTYPE_ENUM:
                raise TypeError(f"{name}: dtype {arr.dtype} cannot be stored in the bundle")
            kvs.append((name.encode(), _entry_proto(arr, offset)))
            f.write(arr.tobytes())
            offset += arr.nbytes
    # data blocks (a new block starts once the current one reaches block_size, as leveldb's TableBuilder does)
    out = bytearray()
    index_kvs, cur, cur_size = [], [], 0

    def flush(next_key):
        nonlocal cur, cur_size
        if not cur:
            return
        block = _build_block(cur)
        handle = _enc_varint(len(out)) + _enc_varint(len(block))
        last = cur[-1][0]
        index_kvs.append((_shortest_separator(last, next_key) if next_key is not None else _short_successor(last), handle))
        out.extend(_with_trailer(block))
        cur, cur_size = [], 0

    for k, v in kvs:
        if cur_size >= block_size:
            flush(k)
        cur.append((k, v))
        cur_size += len(k) + len(v) + 3
    flush(None)
    meta = _build_block([])
    meta_handle = _enc_varint(len(out)) + _enc_varint(len(meta))
    out.extend(_with_trailer(meta))
    index = _build_block_interval1(index_kvs)
    index_handle = _enc_varint(len(out)) + _enc_varint(len(index))
    out.extend(_with_trailer(index))
    footer = meta_handle + index_handle
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC)
    out.extend(footer)
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))


def _build_block_interval1(kvs) -> bytes:
    """Index block: leveldb uses restart interval 1 (every key stored whole)."""
    out, restarts = bytearray(), []
    for k, v in kvs:
        restarts.append(len(out))
        out += _enc_varint(0) + _enc_varint(len(k)) + _enc_varint(len(v)) + k + v
    if not restarts:
        restarts = [0]
    out += struct.pack(f"<{len(restarts)}I", *restarts) + struct.pack("<I", len(restarts))
    return bytes(out)


def load_all(prefix: str) -> Dict[str, np.ndarray]:
    """Every tensor of a bundle (weights, Adam slots, global_step, beta powers), name -> array."""
    out = {}
    with open(prefix + ".data-00000-of-00001", "rb") as f:
        for name, (dtype, shape, off, size) in list_variables(prefix).items():
            f.seek(off)
            out[name] = np.frombuffer(f.read(size), dtype=dtype).reshape(shape).copy()
    return out
