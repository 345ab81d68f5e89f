"""On-disk formats either side of the hot path.

`.flo` (Middlebury): float32 magic 202021.25, int32 width, int32 height, then h*w*2 float32 (u, v interleaved)
-- `load_flow` / `save_flow` follow the reference's flow_utils.py:12-29 byte for byte.
`factor_crop` is test.py:13-17 (crop H and W down to multiples of 64)."""
from __future__ import annotations

import numpy as np

FLO_MAGIC = 202021.25


def load_flow(path):
    """flow_utils.py:12-21: (h, w, 2) float32, or None when the magic number does not match."""
    with open(path, 'rb') as f:
        magic = float(np.fromfile(f, np.float32, count=1)[0])
        if magic != FLO_MAGIC:
            return None
        w = int(np.fromfile(f, np.int32, count=1)[0])
        h = int(np.fromfile(f, np.int32, count=1)[0])
        data = np.fromfile(f, np.float32, count=h * w * 2)
        if data.size != h * w * 2:
            raise ValueError(f"{path}: truncated .flo file ({data.size} of {h * w * 2} floats)")
        return data.reshape(h, w, 2)


def save_flow(path, flow):
    """flow_utils.py:23-29."""
    flow = np.ascontiguousarray(flow, dtype=np.float32)
    if flow.ndim != 3 or flow.shape[2] != 2:
        raise ValueError(f"flow must be (h, w, 2), got {flow.shape}")
    h, w = flow.shape[:2]
    with open(path, 'wb') as f:
        np.array([FLO_MAGIC], np.float32).tofile(f)
        np.array([w], np.int32).tofile(f)
        np.array([h], np.int32).tofile(f)
        flow.tofile(f)


def factor_crop(image, factor=64):
    """test.py:13-17."""
    assert image.ndim == 3
    h, w, _ = image.shape
    return image[:factor * (h // factor), :factor * (w // factor)]
