"""Minimal interpreter for the TensorFlow-1.8 GraphDef the reference saved next to its checkpoints
(`*.ckpt.meta`)  --  TEST INFRASTRUCTURE ONLY (same rules as pwc_oracle.py).

Purpose: pin `oracle/pwc_oracle.py` to the reference's OWN serialized computation.  TensorFlow cannot be
installed here, so the reference program cannot run; but its MetaGraphDef holds the exact forward graph
(`pwcdcnet/*`: 5.8k nodes, 30 op types) and loss graph the reference built from model.py / modules.py /
losses.py, with every attribute (padding, strides, dilation via SpaceToBatchND, slice ranges of the
cost volume, clip bounds, align_corners, concat axes/order, the 0.625..20 scale constants).  This module
parses that protobuf without TensorFlow and executes it node by node with small numpy/torch kernels that
follow the documented TF-1.8 op semantics.  `oracle/make_golden.py` runs it on the trained checkpoint
weights and stores the outputs as fixtures; tests compare the oracle against them.

Only the ops that occur in the reference's forward + loss graph are implemented.
"""
from __future__ import annotations

import struct
from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F

_DT = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 10: np.bool_}


def _varint(buf, pos):
    out, shift = 0, 0
    while True:
        b = buf[pos]; pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf) -> Dict[int, list]:
    out: Dict[int, list] = {}
    pos = 0
    n = len(buf)
    while pos < n:
        tag, pos = _varint(buf, pos)
        f, wt = tag >> 3, tag & 7
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
            raise ValueError(f"wire type {wt}")
        out.setdefault(f, []).append(v)
    return out


def _signed(v, bits=64):
    return v - (1 << bits) if v >= 1 << (bits - 1) else v


def _packed_varints(vals) -> List[int]:
    out = []
    for v in vals:
        if isinstance(v, (bytes, bytearray)):
            pos = 0
            while pos < len(v):
                x, pos = _varint(v, pos)
                out.append(_signed(x))
        else:
            out.append(_signed(v))
    return out


def _shape(buf) -> List[int]:
    return [_signed(_fields(d).get(1, [0])[0]) for d in _fields(buf).get(2, [])]


def _tensor(buf) -> np.ndarray:
    f = _fields(buf)
    dt = _DT[f[1][0]]
    shape = _shape(f[2][0]) if 2 in f else []
    n = int(np.prod(shape)) if shape else 1
    if 4 in f and len(f[4][0]):
        arr = np.frombuffer(f[4][0], dtype=dt).copy()
    elif 5 in f:   # float_val
        vals = []
        for v in f[5]:
            vals += list(struct.unpack(f"<{len(v) // 4}f", v)) if len(v) != 4 else [struct.unpack("<f", v)[0]]
        arr = np.array(vals, dtype=dt)
    elif 7 in f:   # int_val
        arr = np.array(_packed_varints(f[7]), dtype=dt)
    elif 10 in f:  # int64_val
        arr = np.array(_packed_varints(f[10]), dtype=dt)
    elif 11 in f:  # bool_val
        arr = np.array(_packed_varints(f[11]), dtype=dt)
    else:
        arr = np.zeros(n, dtype=dt)
    if arr.size == 1 and n > 1:
        arr = np.full(n, arr[0], dtype=dt)
    return arr.reshape(shape)


def _attr(buf):
    f = _fields(buf)
    if 8 in f:
        return _tensor(f[8][0])
    if 1 in f:   # list
        l = _fields(f[1][0])
        if 3 in l:
            return _packed_varints(l[3])
        if 2 in l:
            return [s for s in l[2]]
        return []
    if 2 in f:
        return f[2][0]
    if 3 in f:
        return _signed(f[3][0])
    if 4 in f:
        return struct.unpack("<f", f[4][0])[0]
    if 5 in f:
        return bool(f[5][0])
    if 6 in f:
        return ("type", f[6][0])
    if 7 in f:
        return _shape(f[7][0])
    return None


class Graph:
    def __init__(self, meta_path: str):
        buf = open(meta_path, "rb").read()
        gd = _fields(_fields(buf)[2][0])
        self.nodes = {}
        for nb in gd[1]:
            n = _fields(nb)
            attrs = {}
            for a in n.get(5, []):
                kv = _fields(a)
                attrs[kv[1][0].decode()] = kv[2][0]
            self.nodes[n[1][0].decode()] = (n[2][0].decode(), [i.decode() for i in n.get(3, [])], attrs)

    def attr(self, name, key):
        a = self.nodes[name][2].get(key)
        return None if a is None else _attr(a)

    # ------------------------------------------------------------------ execution
    def run(self, fetches: List[str], feeds: Dict[str, np.ndarray], variables: Dict[str, np.ndarray]):
        cache: Dict[str, list] = {k: [np.asarray(v)] for k, v in feeds.items()}

        def get(ref):
            ref = ref.lstrip("^")
            name, _, idx = ref.partition(":")
            if name not in cache:
                cache[name] = self._eval(name, get, variables)
            return cache[name][int(idx) if idx else 0]

        import sys
        old = sys.getrecursionlimit()
        sys.setrecursionlimit(20000)
        try:
            return [get(f) for f in fetches]
        finally:
            sys.setrecursionlimit(old)

    def _eval(self, name, get, variables):
        op, inputs, _ = self.nodes[name]
        inputs = [i for i in inputs if not i.startswith("^")]
        x = lambda k: get(inputs[k])
        A = lambda key: self.attr(name, key)
        if op == "Const":
            return [A("value")]
        if op in ("VariableV2", "Variable"):
            return [variables[name]]
        if op in ("Identity", "StopGradient"):
            return [x(0)]
        if op == "Placeholder":
            raise KeyError(f"placeholder {name} must be fed")
        if op == "Conv2D":
            assert A("data_format") in (None, b"NHWC")
            strides, pad = A("strides"), A("padding")
            dil = A("dilations") or [1, 1, 1, 1]
            assert dil == [1, 1, 1, 1], "dilation is expressed with SpaceToBatchND in TF 1.8 graphs"
            inp, k = torch.from_numpy(x(0)), torch.from_numpy(x(1))
            s = strides[1]
            H, W = inp.shape[1], inp.shape[2]
            if pad == b"SAME":
                def p(n):
                    o = -(-n // s)
                    t = max((o - 1) * s + k.shape[0] - n, 0)
                    return t // 2, t - t // 2
                (pt, pb), (pl, pr) = p(H), p(W)
            else:
                pt = pb = pl = pr = 0
            xn = F.pad(inp.permute(0, 3, 1, 2), (pl, pr, pt, pb))
            y = F.conv2d(xn, k.permute(3, 2, 0, 1), None, stride=s)
            return [y.permute(0, 2, 3, 1).contiguous().numpy()]
        if op == "BiasAdd":
            return [x(0) + x(1)]
        if op in ("Mul", "Add", "Sub", "Maximum", "Minimum", "RealDiv", "Less", "GreaterEqual", "FloorDiv", "AddN"):
            a = x(0)
            if op == "AddN":
                out = a
                for k in range(1, len(inputs)):
                    out = out + x(k)
                return [out]
            b = x(1)
            fn = {"Mul": np.multiply, "Add": np.add, "Sub": np.subtract, "Maximum": np.maximum, "Minimum": np.minimum,
                  "RealDiv": np.divide, "Less": np.less, "GreaterEqual": np.greater_equal, "FloorDiv": np.floor_divide}[op]
            r = fn(a, b)
            if op in ("Less", "GreaterEqual"):
                return [r]
            return [r.astype(np.result_type(a, b), copy=False)]
        if op == "Pad":
            pads = x(1)
            return [np.pad(x(0), [(int(a), int(b)) for a, b in pads])]
        if op == "StridedSlice":
            return [self._strided_slice(name, x(0), x(1), x(2), x(3))]
        if op in ("Mean", "Sum"):
            axes = tuple(int(a) for a in np.atleast_1d(x(1)))
            keep = bool(A("keep_dims"))
            fn = np.mean if op == "Mean" else np.sum
            return [fn(x(0), axis=axes, keepdims=keep).astype(x(0).dtype)]
        if op == "Sqrt":
            return [np.sqrt(x(0))]
        if op == "Squeeze":
            dims = A("squeeze_dims")
            return [np.squeeze(x(0), axis=tuple(dims) if dims else None)]
        if op == "L2Loss":
            return [np.array(np.sum(np.square(x(0).astype(np.float64))) / 2, dtype=np.float32)]
        if op == "Square":
            return [np.square(x(0))]
        if op == "Floor":
            return [np.floor(x(0))]
        if op == "Cast":
            dst = _DT[A("DstT")[1]]
            v = x(0)
            if np.issubdtype(dst, np.integer) and np.issubdtype(v.dtype, np.floating):
                v = np.trunc(v)
            return [v.astype(dst)]
        if op == "Pack":
            return [np.stack([x(k) for k in range(len(inputs))], axis=A("axis") or 0)]
        if op == "Unpack":
            ax = A("axis") or 0
            v = x(0)
            return [np.take(v, i, axis=ax) for i in range(v.shape[ax])]
        if op == "Shape":
            return [np.array(x(0).shape, dtype=np.int32)]
        if op == "Size":
            return [np.array(x(0).size, dtype=np.int32)]
        if op == "ClipByValue":
            return [np.minimum(np.maximum(x(0), x(1)), x(2))]
        if op == "GatherNd":
            params, idx = x(0), x(1)
            return [params[tuple(idx[..., k] for k in range(idx.shape[-1]))]]
        if op == "ExpandDims":
            return [np.expand_dims(x(0), int(x(1)))]
        if op == "ConcatV2":
            return [np.concatenate([x(k) for k in range(len(inputs) - 1)], axis=int(x(len(inputs) - 1)))]
        if op == "Range":
            return [np.arange(x(0), x(1), x(2)).astype(np.asarray(x(0)).dtype)]
        if op == "Reshape":
            return [np.reshape(x(0), [int(v) for v in x(1)])]
        if op == "Fill":
            return [np.full([int(v) for v in np.atleast_1d(x(0))], x(1))]
        if op == "Tile":
            return [np.tile(x(0), [int(v) for v in x(1)])]
        if op in ("ResizeBilinear", "ResizeNearestNeighbor"):
            assert not A("align_corners")
            oh, ow = (int(v) for v in x(1))
            return [self._resize(x(0), oh, ow, op == "ResizeBilinear")]
        if op == "SpaceToBatchND":
            return [self._space_to_batch(x(0), x(1), x(2))]
        if op == "BatchToSpaceND":
            return [self._batch_to_space(x(0), x(1), x(2))]
        raise NotImplementedError(f"op {op} (node {name})")

    def _strided_slice(self, name, v, begin, end, strides):
        A = lambda key: self.attr(name, key) or 0
        bm, em, sm, nm, el = A("begin_mask"), A("end_mask"), A("shrink_axis_mask"), A("new_axis_mask"), A("ellipsis_mask")
        assert nm == 0 and el == 0, "new_axis/ellipsis masks do not occur in the reference graph"
        idx = []
        for d in range(len(begin)):
            b, e, s = int(begin[d]), int(end[d]), int(strides[d])
            if sm >> d & 1:
                idx.append(b)
                continue
            idx.append(slice(None if bm >> d & 1 else b, None if em >> d & 1 else e, s))
        return v[tuple(idx)]

    @staticmethod
    def _resize(v, oh, ow, bilinear):
        """TF-1.8 resize kernels, align_corners=False: scale = in/out (float32), src = dst * scale."""
        B, h, w, C = v.shape
        def grid(n_in, n_out):
            scale = np.float32(n_in) / np.float32(n_out)
            s = (np.arange(n_out, dtype=np.float32) * scale).astype(np.float32)
            lo = np.floor(s).astype(np.int64)
            return s, lo
        sy, ylo = grid(h, oh)
        sx, xlo = grid(w, ow)
        if not bilinear:
            return v[:, np.minimum(ylo, h - 1)][:, :, np.minimum(xlo, w - 1)]
        yhi, xhi = np.minimum(ylo + 1, h - 1), np.minimum(xlo + 1, w - 1)
        yl = (sy - ylo.astype(np.float32))[None, :, None, None]
        xl = (sx - xlo.astype(np.float32))[None, None, :, None]
        top, bot = v[:, ylo], v[:, yhi]
        t = top[:, :, xlo] + (top[:, :, xhi] - top[:, :, xlo]) * xl
        b = bot[:, :, xlo] + (bot[:, :, xhi] - bot[:, :, xlo]) * xl
        return (t + (b - t) * yl).astype(np.float32)

    @staticmethod
    def _space_to_batch(v, block, pads):
        by, bx = int(block[0]), int(block[1])
        v = np.pad(v, [(0, 0), (int(pads[0][0]), int(pads[0][1])), (int(pads[1][0]), int(pads[1][1])), (0, 0)])
        B, H, W, C = v.shape
        v = v.reshape(B, H // by, by, W // bx, bx, C).transpose(2, 4, 0, 1, 3, 5)
        return v.reshape(by * bx * B, H // by, W // bx, C)

    @staticmethod
    def _batch_to_space(v, block, crops):
        by, bx = int(block[0]), int(block[1])
        BB, H, W, C = v.shape
        B = BB // (by * bx)
        v = v.reshape(by, bx, B, H, W, C).transpose(2, 3, 0, 4, 1, 5).reshape(B, H * by, W * bx, C)
        return v[:, int(crops[0][0]):H * by - int(crops[0][1]), int(crops[1][0]):W * bx - int(crops[1][1])]


# names of the forward outputs in the reference graph (model.py:95-134 as serialized)
FLOWS_FINAL = "pwcdcnet/mul_6"
MULTISCALE_LOSS = "multiscale_loss/add_4"   # losses.py:15-31 with train.py's weights
TOTAL_LOSS = "add"                           # + gamma * sum l2_loss(var)   (train.py:74-75)
EPE = "Mean"                                 # losses.py:11-13 on flows_final
FLOWS_PYRAMID = ["pwcdcnet/optflow_0/conv2d_5/BiasAdd", "pwcdcnet/optflow_1/add", "pwcdcnet/optflow_2/add",
                 "pwcdcnet/optflow_3/add", "pwcdcnet/context/add"]


def run_reference_graph(meta_path: str, variables: Dict[str, np.ndarray], images: np.ndarray, extra_fetches=(), flows_gt=None):
    """images: (B,2,H,W,3) float32 fed to the `images` placeholder (train.py:46-48)."""
    g = Graph(meta_path)
    feeds = {"images": images.astype(np.float32)}
    if flows_gt is not None:
        feeds["flows"] = flows_gt.astype(np.float32)
    outs = g.run([FLOWS_FINAL] + FLOWS_PYRAMID + list(extra_fetches), feeds, variables)
    return outs[0], outs[1:6], outs[6:]
