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
