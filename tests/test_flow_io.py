"""`.flo` round trip and the reference's factor_crop (flow_utils.py:12-29, test.py:13-17) -- CPU only."""
import numpy as np
import pytest

from pwcnet_b200.flow_io import FLO_MAGIC, factor_crop, load_flow, save_flow


def test_flo_roundtrip_and_byte_layout(tmp_path):
    flow = np.random.default_rng(0).normal(0, 5, (7, 13, 2)).astype(np.float32)
    p = str(tmp_path / "a.flo")
    save_flow(p, flow)
    raw = open(p, "rb").read()
    assert len(raw) == 12 + 7 * 13 * 2 * 4
    assert np.frombuffer(raw[:4], np.float32)[0] == np.float32(FLO_MAGIC)
    assert tuple(np.frombuffer(raw[4:12], np.int32)) == (13, 7)          # width first, then height
    assert np.array_equal(np.frombuffer(raw[12:], np.float32).reshape(7, 13, 2), flow)
    assert np.array_equal(load_flow(p), flow)


def test_flo_bad_magic_and_truncation(tmp_path):
    p = str(tmp_path / "b.flo")
    open(p, "wb").write(np.array([1.0], np.float32).tobytes() + b"\0" * 8)
    assert load_flow(p) is None                                          # flow_utils.py:21
    save_flow(p, np.zeros((4, 4, 2), np.float32))
    open(p, "r+b").truncate(40)
    with pytest.raises(ValueError):
        load_flow(p)
    with pytest.raises(ValueError):
        save_flow(p, np.zeros((4, 4, 3), np.float32))


def test_factor_crop():
    assert factor_crop(np.zeros((436, 1024, 3))).shape == (384, 1024, 3)   # Sintel
    assert factor_crop(np.zeros((448, 1030, 3))).shape == (448, 1024, 3)
    assert factor_crop(np.zeros((63, 64, 3))).shape == (0, 64, 3)
