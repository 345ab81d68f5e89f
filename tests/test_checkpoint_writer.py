"""TF-bundle checkpoint writer (pwcnet_b200/checkpoint.py:save_checkpoint; reference: tf.train.Saver.save at train.py:95,166).
CPU-only: the format is pinned byte for byte against the reference's own checkpoint when /root/reference is mounted."""
import os

import numpy as np
import pytest

from pwcnet_b200 import checkpoint as ck

REF = "/root/reference/model_250epochs_ft_Final/model_250.ckpt"


def test_crc32c_known_answers():
    assert ck._crc32c_py(b"123456789") == 0xE3069283            # the CRC-32C check value
    assert ck._crc32c_py(b"") == 0
    assert ck._crc32c_py(bytes(32)) == 0x8A9136AA                # RFC 3720 B.4: 32 bytes of zeros
    assert ck._crc32c_py(bytes([0xFF] * 32)) == 0x62A8AB43       # RFC 3720 B.4: 32 bytes of ones
    assert ck._crc32c_py(bytes(range(32))) == 0x46DD794E         # RFC 3720 B.4: incrementing bytes
    data = np.random.default_rng(0).integers(0, 256, 100003, dtype=np.uint8).tobytes()
    if os.path.exists(os.path.join(os.path.dirname(ck.__file__), "lib", "libpwc_b200.so")):
        assert ck.crc32c(data) == ck._crc32c_py(data)            # pwc_crc32c (slice-by-8 in the C-ABI library)
        assert ck.crc32c(data[3:70001]) == ck._crc32c_py(data[3:70001])


def test_roundtrip_shapes_dtypes_and_many_blocks(tmp_path):
    rng = np.random.default_rng(1)
    T = {"pwcdcnet/a/kernel": rng.standard_normal((3, 3, 5, 7)).astype(np.float32),
         "pwcdcnet/a/bias": rng.standard_normal(7).astype(np.float32),
         "pwcdcnet/a/kernel/Adam": rng.standard_normal((3, 3, 5, 7)).astype(np.float32),
         "Variable": np.array(41, np.int32), "beta1_power": np.array(0.5, np.float32),
         "count64": np.arange(6, dtype=np.int64).reshape(2, 3), "dbl": np.array([1.5, -2.5])}
    for i in range(300):                                         # enough entries for several table blocks at block_size 2048
        T[f"pwcdcnet/z/var_{i:03d}"] = np.full((i % 3 + 1,), float(i), np.float32)
    prefix = str(tmp_path / "sub" / "model_1.ckpt")
    ck.save_checkpoint(prefix, T, block_size=2048)
    back = ck.load_all(prefix)
    assert set(back) == set(T)
    for k in T:
        assert back[k].dtype == T[k].dtype and back[k].shape == T[k].shape, k
        np.testing.assert_array_equal(back[k], T[k])
    W = ck.load_checkpoint(prefix)                                # the model-variable view (no slots, name filter)
    assert "pwcdcnet/a/kernel" in W and "pwcdcnet/a/kernel/Adam" not in W and "Variable" not in W
    assert int(ck.read_scalar(prefix, "Variable")) == 41


@pytest.mark.skipif(not os.path.exists(REF + ".index"), reason="reference checkpoints not mounted (only in the build container)")
def test_rewriting_the_reference_checkpoint_is_byte_identical(tmp_path):
    """Read every tensor of the reference's model_250 bundle and write it back: .index (entry protos, masked CRC32Cs,
    block layout, footer) and .data must equal the files tf.train.Saver produced, byte for byte."""
    T = ck.load_all(REF)
    assert len(T) == 333
    prefix = str(tmp_path / "model_250.ckpt")
    ck.save_checkpoint(prefix, T)
    for sfx in (".index", ".data-00000-of-00001"):
        with open(REF + sfx, "rb") as a, open(prefix + sfx, "rb") as b:
            assert a.read() == b.read(), sfx
