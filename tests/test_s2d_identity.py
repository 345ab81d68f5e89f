"""CPU proof of the identity the stride-2 halo path relies on (csrc/conv_tc_f16.cu, pwc_conv3x3_s2_tc_f16_fwd): a 3x3 stride-2
TF-'SAME' convolution of an even-sized image equals a 2x2 stride-1 convolution over the space-to-depth view
X'[y][x][(py,px,c)] = X[2y+py][2x+px][c] with the re-indexed kernel W'[dy][dx][(py,px,c)] = W[2dy+py][2dx+px][c] (zero where
2dy+py or 2dx+px would be 3) and zero padding AFTER the last row / column only.  The GPU test
(tests/test_gpu_ops.py::test_conv_stride2_space_to_depth_matches_oracle) checks the kernel; this one checks the algebra against
the oracle's conv (modules.py:62-63 semantics) without a GPU."""
import numpy as np
import pytest
import torch

from oracle import pwc_oracle as O


def s2d_reindex_ref(k):
    """numpy restatement of pwc_conv3x3_s2d_reindex: (3,3,C,Co) -> (3,3,4C,Co), taps (0..1, 0..1) used."""
    C, Co = k.shape[2], k.shape[3]
    out = np.zeros((3, 3, 2, 2, C, Co), np.float32)
    for dy in range(2):
        for dx in range(2):
            for py in range(2):
                for px in range(2):
                    ky, kx = 2 * dy + py, 2 * dx + px
                    if ky < 3 and kx < 3:
                        out[dy, dx, py, px] = k[ky, kx]
    return out.reshape(3, 3, 4 * C, Co)


@pytest.mark.parametrize("shape,cout", [((2, 8, 12, 16), 32), ((1, 6, 2, 5), 3), ((1, 2, 2, 4), 4), ((3, 14, 10, 7), 9)])
def test_stride2_conv_is_a_2x2_conv_over_space_to_depth(shape, cout):
    B, H, W, C = shape
    rng = np.random.default_rng(0)
    x = rng.standard_normal(shape).astype(np.float32)
    k = (rng.standard_normal((3, 3, C, cout)) * 0.1).astype(np.float32)
    b = rng.standard_normal((cout,)).astype(np.float32)
    ref = O.conv2d_same(torch.from_numpy(x), torch.from_numpy(k), torch.from_numpy(b), stride=2).numpy()
    k2 = s2d_reindex_ref(k)
    # space-to-depth view, zero-padded by one row / column after the image
    xs = x.reshape(B, H // 2, 2, W // 2, 2, C).transpose(0, 1, 3, 2, 4, 5).reshape(B, H // 2, W // 2, 4 * C)
    xp = np.zeros((B, H // 2 + 1, W // 2 + 1, 4 * C), np.float64)
    xp[:, :-1, :-1] = xs
    y = np.zeros((B, H // 2, W // 2, cout), np.float64)
    for dy in range(2):
        for dx in range(2):
            y += np.einsum("bhwc,co->bhwo", xp[:, dy:dy + H // 2, dx:dx + W // 2], k2[dy, dx].astype(np.float64))
    y += b
    np.testing.assert_allclose(y, ref, rtol=0, atol=2e-5)
    # the unused taps of the re-indexed kernel are exactly zero
    assert not k2[2].any() and not k2[:, 2].any()
