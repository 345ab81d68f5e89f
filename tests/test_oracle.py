"""Known-answer tests that pin the CPU oracle to the TF-1.8 semantics the reference relies on
(SURVEY.md section 9).  No GPU."""
import os

import numpy as np
import pytest
import torch

from oracle import pwc_oracle as O

REF = "/root/reference"


def test_same_padding_tf_asymmetric():
    assert O.same_padding(8, 3, 2, 1) == (0, 1)      # even size, stride 2: 0 before / 1 after
    assert O.same_padding(7, 3, 2, 1) == (1, 1)
    assert O.same_padding(448, 3, 2, 1) == (0, 1)
    assert O.same_padding(9, 3, 1, 1) == (1, 1)
    assert O.same_padding(9, 3, 1, 16) == (16, 16)


def test_conv_stride2_pads_after_not_before():
    x = torch.arange(16, dtype=torch.float32).view(1, 4, 4, 1)
    k = torch.ones(3, 3, 1, 1)
    y = O.conv2d_same(x, k, torch.zeros(1), stride=2)
    assert y.shape == (1, 2, 2, 1)
    # window of out[0,0] starts at input (0,0) (no top/left pad): rows 0..2, cols 0..2
    assert y[0, 0, 0, 0].item() == x[0, 0:3, 0:3, 0].sum().item()
    # out[1,1] window starts at (2,2), one zero row/col after the image
    assert y[0, 1, 1, 0].item() == x[0, 2:4, 2:4, 0].sum().item()


def test_conv_is_cross_correlation_hwio_with_bias():
    x = torch.zeros(1, 5, 5, 2); x[0, 2, 2, 1] = 1.0
    k = torch.arange(3 * 3 * 2 * 3, dtype=torch.float32).view(3, 3, 2, 3)
    b = torch.tensor([10.0, 20.0, 30.0])
    y = O.conv2d_same(x, k, b)
    # impulse at (2,2) channel 1: y[2-dy+1, 2-dx+1] = k[dy,dx,1,:] + b   (no kernel flip)
    for dy in range(3):
        for dx in range(3):
            np.testing.assert_allclose(y[0, 3 - dy, 3 - dx].numpy(), (k[dy, dx, 1] + b).numpy())


def test_dilated_conv_same():
    x = torch.zeros(1, 9, 9, 1); x[0, 4, 4, 0] = 1.0
    k = torch.arange(9, dtype=torch.float32).view(3, 3, 1, 1) + 1
    y = O.conv2d_same(x, k, torch.zeros(1), dilation=2)
    assert y.shape == x.shape
    for dy in range(3):
        for dx in range(3):
            assert y[0, 4 - 2 * (dy - 1), 4 - 2 * (dx - 1), 0].item() == k[dy, dx, 0, 0].item()


def test_leaky_relu_is_max():
    x = torch.tensor([-2.0, 0.0, 3.0])
    np.testing.assert_allclose(O.leaky_relu(x, 0.1).numpy(), [-0.2, 0.0, 3.0], rtol=1e-7)


def test_resize_bilinear_legacy_x2_x4_rows():
    v = torch.tensor([1.0, 3.0, 7.0, 4.0]).view(1, 1, 4, 1)
    y = O.resize_bilinear_legacy(v, 1, 8).flatten().numpy()
    # out[2k] = in[k]; out[2k+1] = (in[k] + in[min(k+1,n-1)])/2   -- no half-pixel offset
    np.testing.assert_allclose(y, [1, 2, 3, 5, 7, 5.5, 4, 4])
    y4 = O.resize_bilinear_legacy(v, 1, 16).flatten().numpy()
    np.testing.assert_allclose(y4[:5], [1, 1.5, 2, 2.5, 3])
    np.testing.assert_allclose(y4[12:], [4, 4, 4, 4])
    # differs from torch's half-pixel bilinear
    t = torch.nn.functional.interpolate(v.permute(0, 3, 1, 2), size=(1, 8), mode="bilinear", align_corners=False)
    assert not np.allclose(t.flatten().numpy(), y)


def test_resize_nearest_is_strided_subsample():
    x = torch.arange(2 * 8 * 12 * 2, dtype=torch.float32).view(2, 8, 12, 2)
    np.testing.assert_array_equal(O.resize_nearest_legacy(x, 2, 3).numpy(), x[:, ::4, ::4].numpy())


def test_bilinear_warp_integer_shift_and_border_replication():
    x = torch.arange(5 * 6, dtype=torch.float32).view(1, 5, 6, 1)
    flow = torch.zeros(1, 5, 6, 2); flow[..., 0] = 2.0; flow[..., 1] = -1.0   # x+2, y-1
    y = O.bilinear_warp(x, flow)
    for yy in range(5):
        for xx in range(6):
            assert y[0, yy, xx, 0] == x[0, max(yy - 1, 0), min(xx + 2, 5), 0]
    # far outside: border value with full weight (weights are not clamped -> sum to 1)
    flow[..., 0] = 100.3
    y = O.bilinear_warp(x, flow)
    np.testing.assert_allclose(y[0, :, :, 0].numpy(), x[0, [0, 0, 1, 2, 3], 5, 0].view(5, 1).expand(5, 6).numpy(), rtol=1e-5)


def test_bilinear_warp_fractional_weights():
    x = torch.tensor([[0.0, 10.0], [100.0, 1000.0]]).view(1, 2, 2, 1)
    flow = torch.zeros(1, 2, 2, 2); flow[0, 0, 0] = torch.tensor([0.25, 0.5])
    y = O.bilinear_warp(x, flow)
    expect = 0.5 * 0.75 * 0 + 0.5 * 0.25 * 10 + 0.5 * 0.75 * 100 + 0.5 * 0.25 * 1000
    np.testing.assert_allclose(y[0, 0, 0, 0].item(), expect, rtol=1e-6)


def test_nearest_warp_truncates_toward_zero():
    x = torch.arange(7, dtype=torch.float32).view(1, 1, 7, 1)
    flow = torch.zeros(1, 1, 7, 2)
    flow[0, 0, :, 0] = torch.tensor([-0.7, 1.7, -1.2, 0.99, -2.0, 5.0, -0.01])
    y = O.nearest_warp(x, flow).flatten().numpy()
    np.testing.assert_array_equal(y, [0, 2, 1, 3, 2, 6, 6])


def test_cost_volume_channel_order_mean_and_leaky():
    C = 4
    f0 = torch.zeros(1, 12, 12, C); f1 = torch.zeros(1, 12, 12, C)
    f0[0, 5, 6, 2] = 2.0
    f1[0, 5 + 3, 6 - 2, 2] = 3.0      # v = +3 (down), h = -2 (left)
    f1[0, 5 - 4, 6 + 4, 2] = -1.0     # v = -4, h = +4  (negative -> leaky)
    cv = O.cost_volume(f0, f1, 4)
    assert cv.shape == (1, 12, 12, 81)
    d1 = (3 + 4) * 9 + (-2 + 4)
    d2 = (-4 + 4) * 9 + (4 + 4)
    np.testing.assert_allclose(cv[0, 5, 6, d1].item(), 2.0 * 3.0 / C)
    np.testing.assert_allclose(cv[0, 5, 6, d2].item(), 0.1 * (2.0 * -1.0 / C), rtol=1e-6)
    nz = torch.nonzero(cv)
    assert len(nz) == 2


def test_cost_volume_structure_equals_closed_form():
    g = torch.Generator().manual_seed(0)
    f0 = torch.randn(2, 9, 11, 8, generator=g); f1 = torch.randn(2, 9, 11, 8, generator=g)
    np.testing.assert_allclose(O.cost_volume(f0, f1).numpy(), O.cost_volume_closed_form(f0, f1).numpy(), atol=1e-6)
    np.testing.assert_allclose(O.cost_volume(f0, f1, 2).numpy(), O.cost_volume_closed_form(f0, f1, 2).numpy(), atol=1e-6)


def test_losses_hand_values():
    gt = torch.zeros(2, 4, 4, 2); fl = torch.zeros(2, 4, 4, 2)
    fl[..., 0] = 3.0; fl[..., 1] = 4.0
    assert O.EPE(gt, fl).item() == pytest.approx(5.0)
    assert O.L2loss(gt, fl).item() == pytest.approx(5.0 * 16)
    assert O.L1loss(gt, fl).item() == pytest.approx(7.0 * 16)
    gt20 = torch.full((2, 8, 8, 2), 20.0)
    pyr = [torch.zeros(2, 2, 2, 2), torch.zeros(2, 4, 4, 2)]
    # ||(1,1)|| = sqrt2 per pixel
    expect = 0.5 * np.sqrt(2) * 4 + 0.25 * np.sqrt(2) * 16
    assert O.multiscale_loss(gt20, pyr, [0.5, 0.25]).item() == pytest.approx(expect, rel=1e-6)


def test_param_count_matches_reference_checkpoints():
    W = O.glorot_weights(2)
    assert len(W) == 110 and sum(v.size for v in W.values()) == 5029868
    Wd = O.glorot_weights(2, use_dc=True)
    assert sum(v.size for v in Wd.values()) == 40182338
    assert W["pwcdcnet/optflow_0/conv2d/kernel"].shape == (3, 3, 273, 128)
    assert W["pwcdcnet/optflow_4/conv2d/kernel"].shape == (3, 3, 147, 128)
    assert W["pwcdcnet/context/conv2d/kernel"].shape == (3, 3, 34, 128)
    lim = np.sqrt(6 / (27 + 144))
    assert abs(W["pwcdcnet/fp_extractor/conv2d/kernel"]).max() <= lim + 1e-7       # +-0.18731716 in the GraphDef


def test_config1_shapes_and_golden():
    """BASELINE config 1: single 2x3x64x128 pair, random-init forward: per-scale shapes + EPE."""
    W = O.glorot_weights(2)
    im0, im1 = O.synthetic_pair(1, 64, 128, 0)
    ff, pyr = O.pwcdcnet_forward(W, im0, im1)
    assert ff.shape == (1, 64, 128, 2)
    assert [tuple(p.shape) for p in pyr] == [(1, 1, 2, 2), (1, 2, 4, 2), (1, 4, 8, 2), (1, 8, 16, 2), (1, 16, 32, 2)]
    # golden = the reference's own saved GraphDef executed node by node (oracle/make_golden.py)
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "config1_glorot_seed2.npz"))
    assert "reference GraphDef" in str(gold["source"])
    np.testing.assert_allclose(ff.numpy(), gold["flows_final"], atol=5e-6)
    for l, p in enumerate(pyr):
        np.testing.assert_allclose(p.numpy(), gold[f"pyr{l}"], atol=2e-6)
    gt = np.random.default_rng(1).normal(0, 5, (1, 64, 128, 2)).astype(np.float32)
    assert O.EPE(torch.from_numpy(gt), ff).item() == pytest.approx(float(gold["epe"]), rel=1e-5)


def test_hot_weights_vs_reference_graph_golden():
    """Flows up to 14 px (warp / clamp / residual paths live): oracle vs the reference GraphDef's outputs,
    including the serialized loss graph (multiscale loss, + 4e-4 * sum l2_loss(var), EPE)."""
    W = O.glorot_weights(7, gain=1.4, bias_scale=0.02)
    im0, im1 = O.synthetic_pair(2, 64, 128, 3, shift=(5, -3))
    gt = np.random.default_rng(1).normal(0, 5, (2, 64, 128, 2)).astype(np.float32)
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "hot_seed7_64x128.npz"))
    loss, epe, ff, pyr = O.training_loss({k: torch.from_numpy(v) for k, v in W.items()}, im0, im1, gt)
    assert np.abs(gold["flows_final"]).max() > 10
    np.testing.assert_allclose(ff.numpy(), gold["flows_final"], atol=2e-4)
    for l, p in enumerate(pyr):
        np.testing.assert_allclose(p.numpy(), gold[f"pyr{l}"], atol=1e-5)
    assert loss.item() == pytest.approx(float(gold["total_loss"]), rel=1e-5)
    assert O.multiscale_loss(torch.from_numpy(gt), pyr, O.DEFAULT_LOSS_WEIGHTS).item() == pytest.approx(float(gold["loss"]), rel=1e-5)
    assert epe.item() == pytest.approx(float(gold["epe"]), rel=1e-5)


@pytest.mark.skipif(not os.path.exists(REF + "/model_250epochs_ft_Final/model_250.ckpt.meta"),
                    reason="reference GraphDef not mounted (only in the build container)")
def test_reference_graphdef_live_with_trained_checkpoint():
    """Execute the reference's saved GraphDef with its trained weights and compare the oracle, live."""
    from oracle import tf_graph_interp as G
    from pwcnet_b200.checkpoint import load_checkpoint
    ck = REF + "/model_250epochs_ft_Final/model_250.ckpt"
    W = load_checkpoint(ck)
    im0, im1 = O.synthetic_pair(1, 64, 128, 0, shift=(3, -2))
    ff, pyr, _ = G.run_reference_graph(ck + ".meta", W, np.stack([im0, im1], 1))
    rff, rpyr = O.pwcdcnet_forward(W, im0, im1)
    np.testing.assert_allclose(rff.numpy(), ff, atol=2e-5)
    for a, b in zip(rpyr, pyr):
        np.testing.assert_allclose(a.numpy(), b, atol=2e-6)
    # graph-level facts the restatement relies on
    g = G.Graph(ck + ".meta")
    assert float(g.attr("pwcdcnet/mul_6/y", "value")) == 20.0                        # model.py:127
    assert [float(g.attr(f"pwcdcnet/mul{'_' + str(i) if i else ''}/y", "value")) for i in range(4)] == [0.625, 1.25, 2.5, 5.0]
    assert g.attr("pwcdcnet/ResizeBilinear", "align_corners") in (None, False)
    assert g.attr("pwcdcnet/fp_extractor/conv2d/Conv2D", "padding") == b"SAME"
    assert g.attr("pwcdcnet/fp_extractor/conv2d/Conv2D", "strides") == [1, 2, 2, 1]


def test_piecewise_lr_and_adam():
    assert O.piecewise_lr(0) == 1e-4 and O.piecewise_lr(200000) == 1e-4
    assert O.piecewise_lr(200001) == 5e-5 and O.piecewise_lr(5000000) == 1e-4 / 32
    var, m, v = O.adam_step_tf(np.ones(3, np.float32), np.full(3, 0.5, np.float32), np.zeros(3, np.float32),
                               np.zeros(3, np.float32), 1, 1e-4)
    # first step: m = .05, v = .00025*..., update = lr * sqrt(1-b2)/(1-b1) * m/(sqrt(v)+eps) ~= lr
    np.testing.assert_allclose(var, 1 - 1e-4, rtol=1e-6)


@pytest.mark.skipif(not os.path.exists(REF + "/model_250epochs_ft_Final/model_250.ckpt.index"),
                    reason="reference checkpoints not mounted (only in the build container)")
def test_trained_checkpoint_recovers_known_translation():
    """End-to-end semantic pin: with the reference's own trained weights, the oracle must recover a
    synthetic translation.  Any wrong TF-1.8 semantic (padding, resize, warp, channel order, /20
    scaling) breaks this."""
    from pwcnet_b200.checkpoint import load_checkpoint
    W = load_checkpoint(REF + "/model_250epochs_ft_Final/model_250.ckpt")
    assert len(W) == 110
    for (H, Wd), tol in (((64, 128), 0.25), ((128, 192), 0.6)):   # white-noise texture: harder at finer scale
        im0, im1 = O.synthetic_pair(1, H, Wd, 0, shift=(3, -2))
        ff, _ = O.pwcdcnet_forward(W, im0, im1)
        med = np.median(ff.numpy().reshape(-1, 2), axis=0)
        assert abs(med[0] - 3.0) < tol and abs(med[1] + 2.0) < tol, med
