"""GPU parity of every C-ABI operator against the CPU oracle on the same seeded inputs.
Tolerances: these are float32 kernels whose only difference from the oracle is summation order
(and FMA contraction), so max-abs 1e-5 on O(1) data; stated per test."""
import os

import numpy as np
import pytest
import torch

from oracle import pwc_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import pwcnet_b200 as P
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    P.ops.lib()
    return P


def _rand(shape, seed, scale=1.0):
    return (np.random.default_rng(seed).standard_normal(shape) * scale).astype(np.float32)


def _cuda(a):
    return torch.from_numpy(a).cuda()


# the five pyramid levels of a 448x1024 pair scaled down, plus ragged / tiny / odd shapes
CV_SHAPES = [(2, 7, 16, 192), (1, 14, 32, 128), (1, 28, 64, 96), (1, 56, 128, 64), (1, 112, 256, 32),
             (2, 5, 9, 16), (1, 1, 2, 32), (3, 13, 45, 36), (1, 9, 40, 4),
             (3, 13, 45, 48), (2, 30, 70, 16), (1, 8, 33, 64), (1, 15, 100, 32)]


@pytest.mark.parametrize("shape", CV_SHAPES)
def test_cost_volume_matches_oracle(P, shape):
    f0, f1 = _rand(shape, 1), _rand(shape, 2)
    ref = O.cost_volume(torch.from_numpy(f0), torch.from_numpy(f1), 4).numpy()
    out = P.ops.cost_volume(_cuda(f0), _cuda(f1), 4).cpu().numpy()
    assert out.shape == ref.shape
    np.testing.assert_allclose(out, ref, atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("shape", [(2, 7, 16, 192), (1, 28, 64, 96), (1, 112, 256, 32), (3, 13, 45, 64), (1, 9, 17, 32), (2, 30, 70, 32)])
@pytest.mark.parametrize("slot", [False, True])
def test_cost_volume_tcgen05_variant_matches_oracle(P, shape, slot, monkeypatch):
    """The opt-in tensor-core band-GEMM kernel (PWC_CV_KERNEL=tc, 3 x fp16 split, fp32-class): dense output and
    the 81-channel slot of a wider concat buffer with the f0 copy, ragged tiles, all five level widths."""
    import os
    monkeypatch.setenv("PWC_CV_KERNEL", "tc")
    os.environ["PWC_CV_KERNEL"] = "tc"
    f0, f1 = _rand(shape, 1), _rand(shape, 2)
    ref = O.cost_volume(torch.from_numpy(f0), torch.from_numpy(f1), 4).numpy()
    B, H, W, C = shape
    if slot:
        buf = torch.full((B, H, W, 84 + C), 7.0, device="cuda")
        out = P.ops.cost_volume(_cuda(f0), _cuda(f1), 4, out=buf[..., :81], f0_copy=buf[..., 84:84 + C])
        np.testing.assert_array_equal(buf[..., 84:].cpu().numpy(), f0)
        assert float((buf[..., 81:84] - 7.0).abs().max()) == 0
    else:
        out = P.ops.cost_volume(_cuda(f0), _cuda(f1), 4)
    np.testing.assert_allclose(out.cpu().numpy(), ref, atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("shape", [(2, 7, 16, 192), (1, 28, 64, 96), (1, 112, 256, 32), (3, 13, 45, 64), (1, 9, 17, 32), (2, 30, 70, 32)])
def test_cost_volume_split_pipeline_matches_oracle(P, shape):
    """The pipeline's fast path: split producers (pwc_split_f16_fwd with the f0 slot copy, pwc_warp_split_fwd) feeding
    the tcgen05 cost volume; result vs the oracle's warp + cost volume on fp32 (1e-5 max-abs: fp32-class)."""
    B, H, W, C = shape
    f0, f1 = _rand(shape, 1), _rand(shape, 2)
    flow = _rand((B, H, W, 2), 3, 1.7)
    ref_w = O.bilinear_warp(torch.from_numpy(f1), torch.from_numpy(flow) * 2.5)
    ref = O.cost_volume(torch.from_numpy(f0), ref_w, 4).numpy()
    buf = torch.full((B, H, W, 84 + C), 7.0, device="cuda")
    f0s = P.ops.split_f16(_cuda(f0), copy=buf[..., 84:84 + C])
    f1s = P.ops.warp_split(_cuda(f1), _cuda(flow), 2.5, "bilinear")
    # the split tensors reproduce their fp32 source to fp32 precision: h + l
    hl = f0s.float().view(B, H, W, C // 32, 2, 32)
    np.testing.assert_allclose((hl[..., 0, :] + hl[..., 1, :]).reshape(B, H, W, C).cpu().numpy(), f0, atol=4e-7, rtol=1e-6)
    np.testing.assert_array_equal(buf[..., 84:].cpu().numpy(), f0)
    P.ops.cost_volume_split(f0s, f1s, 0.1, out=buf[..., :81])
    np.testing.assert_allclose(buf[..., :81].cpu().numpy(), ref, atol=1e-5, rtol=1e-5)
    assert float((buf[..., 81:84] - 7.0).abs().max()) == 0
    dense = P.ops.cost_volume_split(f0s, f1s, 0.1)
    np.testing.assert_allclose(dense.cpu().numpy(), ref, atol=1e-5, rtol=1e-5)
    # the pipeline's form: 1/C folded into the f0 producer, the slot copy stays unscaled
    buf2 = torch.full((B, H, W, 84 + C), 7.0, device="cuda")
    f0p = P.ops.split_f16(_cuda(f0), copy=buf2[..., 84:84 + C], scale=1.0 / C)
    np.testing.assert_array_equal(buf2[..., 84:].cpu().numpy(), f0)
    P.ops.cost_volume_split(f0p, f1s, 0.1, out=buf2[..., :81], prescaled=True)
    np.testing.assert_allclose(buf2[..., :81].cpu().numpy(), ref, atol=1e-5, rtol=1e-5)
    f1n = P.ops.warp_split(_cuda(f1), _cuda(flow), 2.5, "nearest")
    refn = O.cost_volume(torch.from_numpy(f0), O.nearest_warp(torch.from_numpy(f1), torch.from_numpy(flow) * 2.5), 4).numpy()
    np.testing.assert_allclose(P.ops.cost_volume_split(f0s, f1n, 0.1).cpu().numpy(), refn, atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("variant", ["quad", "scatter"])
@pytest.mark.parametrize("shape", [(1, 8, 64, 32), (2, 13, 70, 32), (1, 28, 64, 64), (1, 5, 31, 96), (1, 33, 9, 32), (2, 16, 8, 32),
                                   (1, 112, 256, 32), (2, 7, 16, 192)])
def test_cost_volume_split_variants(P, shape, variant, monkeypatch):
    """Both tilings of the split band GEMM -- the quadrant-block kernel (default: 16 x 8 tiles, register-resident band
    extraction) and the round-1 scatter kernel (PWC_CV_SPLIT=scatter) -- vs the oracle, into a strided slot, on ragged
    shapes (partial tiles right and bottom, images narrower / shorter than one tile)."""
    B, H, W, C = shape
    f0, f1 = _rand(shape, 1), _rand(shape, 2)
    ref = O.cost_volume_closed_form(torch.from_numpy(f0), torch.from_numpy(f1), 4).numpy() if H * W > 4000 else \
        O.cost_volume(torch.from_numpy(f0), torch.from_numpy(f1), 4).numpy()
    f0s, f1s = P.ops.split_f16(_cuda(f0)), P.ops.split_f16(_cuda(f1))
    monkeypatch.setenv("PWC_CV_SPLIT", variant)
    buf = torch.full((B, H, W, 88), 7.0, device="cuda")
    P.ops.cost_volume_split(f0s, f1s, 0.1, out=buf[..., :81])
    np.testing.assert_allclose(buf[..., :81].cpu().numpy(), ref, atol=1e-5, rtol=1e-5)
    assert float((buf[..., 81:] - 7.0).abs().max()) == 0
    # unaligned destination (scalar store path): channel offset 1 inside the buffer
    buf2 = torch.full((B, H, W, 88), 7.0, device="cuda")
    P.ops.cost_volume_split(f0s, f1s, 0.1, out=buf2[..., 1:82])
    np.testing.assert_array_equal(buf2[..., 1:82].cpu().numpy(), buf[..., :81].cpu().numpy())
    assert float((buf2[..., 82:] - 7.0).abs().max()) == 0 and float((buf2[..., :1] - 7.0).abs().max()) == 0


@pytest.mark.parametrize("head", [88])
@pytest.mark.parametrize("shape", [(2, 13, 70, 32), (1, 112, 256, 32), (2, 7, 16, 192), (1, 33, 9, 64)])
def test_cost_volume_split_slot_mode_writes_whole_sectors(P, shape, head):
    """pwc_cost_volume_split_slot_fwd: words [0,81) = cost volume, [81,83) = the tail tensor (up-sampled flow), [83,88) = 0,
    everything beyond word 88 untouched; interior and ragged tiles."""
    B, H, W, C = shape
    f0, f1 = _rand(shape, 1), _rand(shape, 2)
    tail = _rand((B, H, W, 2), 3)
    ref = O.cost_volume_closed_form(torch.from_numpy(f0), torch.from_numpy(f1), 4).numpy()
    f0s, f1s = P.ops.split_f16(_cuda(f0)), P.ops.split_f16(_cuda(f1))
    buf = torch.full((B, H, W, 96 + C), 7.0, device="cuda")
    P.ops.cost_volume_split(f0s, f1s, 0.1, out=buf[..., :81], slot=head, tail=_cuda(tail))
    np.testing.assert_allclose(buf[..., :81].cpu().numpy(), ref, atol=1e-5, rtol=1e-5)
    np.testing.assert_array_equal(buf[..., 81:83].cpu().numpy(), tail)
    assert float(buf[..., 83:head].abs().max()) == 0 and float((buf[..., head:] - 7.0).abs().max()) == 0
    buf.fill_(7.0)
    P.ops.cost_volume_split(f0s, f1s, 0.1, out=buf[..., :81], slot=head)
    np.testing.assert_allclose(buf[..., :81].cpu().numpy(), ref, atol=1e-5, rtol=1e-5)
    assert float(buf[..., 81:head].abs().max()) == 0 and float((buf[..., head:] - 7.0).abs().max()) == 0
    with pytest.raises(P.PwcError):
        P.ops.cost_volume_split(f0s, f1s, 0.1, out=torch.empty((B, H, W, 84), device="cuda")[..., :81], slot=head)


def test_cost_volume_split_all_81_displacements_at_corners(P):
    """Impulses in the four corners through the tcgen05 split path: every displacement channel picks exactly the
    reference's (v outer, h inner) neighbour and zero-pads outside the image."""
    C, H, W = 32, 19, 35
    f0 = np.zeros((1, H, W, C), np.float32); f1 = _rand((1, H, W, C), 5)
    for (y, x) in [(0, 0), (0, W - 1), (H - 1, 0), (H - 1, W - 1), (9, 17)]:
        f0[0, y, x, :4] = [1.0, -2.0, 0.5, 3.0]
    ref = O.cost_volume(torch.from_numpy(f0), torch.from_numpy(f1), 4).numpy()
    out = P.ops.cost_volume_split(P.ops.split_f16(_cuda(f0)), P.ops.split_f16(_cuda(f1)), 0.1).cpu().numpy()
    np.testing.assert_allclose(out, ref, atol=1e-6)
    assert (out[0, 0, 0, :4 * 9] == 0).all()


def test_cost_volume_all_81_displacements_at_corners(P):
    """Adversarial: impulses in the four corners; every displacement channel must pick exactly the
    reference's (v outer, h inner) neighbour and zero-pad outside the image."""
    C, H, W = 4, 9, 35
    f0 = np.zeros((1, H, W, C), np.float32); f1 = _rand((1, H, W, C), 5)
    for (y, x) in [(0, 0), (0, W - 1), (H - 1, 0), (H - 1, W - 1)]:
        f0[0, y, x, :] = [1.0, -2.0, 0.5, 3.0]
    ref = O.cost_volume(torch.from_numpy(f0), torch.from_numpy(f1), 4).numpy()
    out = P.ops.cost_volume(_cuda(f0), _cuda(f1), 4).cpu().numpy()
    np.testing.assert_allclose(out, ref, atol=1e-6)
    assert (out[0, 0, 0, :4 * 9] == 0).all()          # v < 0 rows are outside the image at y = 0


@pytest.mark.parametrize("r", [1, 2, 3])
def test_cost_volume_other_search_ranges(P, r):
    f0, f1 = _rand((2, 10, 12, 8), 1), _rand((2, 10, 12, 8), 2)
    ref = O.cost_volume(torch.from_numpy(f0), torch.from_numpy(f1), r).numpy()
    out = P.ops.cost_volume(_cuda(f0), _cuda(f1), r).cpu().numpy()
    np.testing.assert_allclose(out, ref, atol=1e-5)


def test_cost_volume_into_concat_slot_with_f0_copy(P):
    B, H, W, C = 2, 14, 40, 32
    f0, f1 = _rand((B, H, W, C), 1), _rand((B, H, W, C), 2)
    buf = torch.full((B, H, W, 148), -7.0, device="cuda")
    P.ops.cost_volume(_cuda(f0), _cuda(f1), 4, out=buf[..., 0:81], f0_copy=buf[..., 84:116])
    ref = O.cost_volume(torch.from_numpy(f0), torch.from_numpy(f1), 4).numpy()
    got = buf.cpu().numpy()
    np.testing.assert_allclose(got[..., 0:81], ref, atol=1e-5)
    np.testing.assert_array_equal(got[..., 84:116], f0)
    assert (got[..., 81:84] == -7.0).all() and (got[..., 116:] == -7.0).all()     # nothing else touched


@pytest.mark.parametrize("warp_type", ["bilinear", "nearest"])
@pytest.mark.parametrize("shape", [(2, 14, 32, 128), (1, 28, 64, 96), (2, 9, 21, 16)])
def test_warp_matches_oracle(P, shape, warp_type):
    B, H, W, C = shape
    x = _rand(shape, 3)
    flow = _rand((B, H, W, 2), 4, scale=6.0)      # many samples beyond the borders
    flow[0, 0, 0] = [0.0, 0.0]; flow[0, 1, 1] = [2.0, -1.0]; flow[0, 2, 2] = [-100.0, 100.0]   # integer / far flows
    ref = O.warping_layer(torch.from_numpy(x), torch.from_numpy(flow), warp_type).numpy()
    out = P.ops.warp(_cuda(x), _cuda(flow), 1.0, warp_type).cpu().numpy()
    np.testing.assert_allclose(out, ref, atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("warp_type", ["bilinear", "nearest"])
def test_fused_warp_cost_volume_equals_warp_then_cost_volume(P, warp_type):
    B, H, W, C = 2, 28, 64, 96
    f0, f1 = _rand((B, H, W, C), 1), _rand((B, H, W, C), 2)
    flow = _rand((B, H, W, 2), 4, scale=3.0)
    scale = 2.5
    fl = torch.from_numpy(flow) * scale
    ref = O.cost_volume(torch.from_numpy(f0), O.warping_layer(torch.from_numpy(f1), fl, warp_type), 4).numpy()
    out = P.ops.warp_cost_volume(_cuda(f0), _cuda(f1), _cuda(flow), scale, warp_type).cpu().numpy()
    np.testing.assert_allclose(out, ref, atol=2e-5, rtol=1e-5)


CONV_CASES = [
    # (B, H, W, Cin, Cout, stride, dilation, alpha)
    (2, 16, 24, 3, 16, 2, 1, 0.1),      # first pyramid conv: Cin=3, stride 2 (asymmetric SAME)
    (1, 15, 17, 3, 16, 2, 1, 0.1),      # odd sizes
    (2, 16, 24, 16, 16, 1, 1, 0.1),
    (1, 14, 32, 16, 32, 2, 1, 0.1),     # stride-2 with Cin >= 16
    (1, 7, 16, 273, 128, 1, 1, 0.1),    # estimator conv 0 at the coarsest level
    (1, 12, 20, 147, 128, 1, 1, 0.1),
    (1, 12, 20, 128, 96, 1, 1, 0.1),
    (2, 12, 20, 32, 2, 1, 1, 1.0),      # flow head
    (1, 24, 40, 34, 128, 1, 1, 0.1),    # context conv 0
    (1, 24, 40, 128, 128, 1, 2, 0.1),
    (1, 24, 40, 128, 128, 1, 4, 0.1),
    (1, 24, 40, 128, 96, 1, 8, 0.1),
    (1, 24, 40, 96, 64, 1, 16, 0.1),    # dilation 16 on a small map: mostly padding
    (1, 9, 11, 5, 7, 1, 1, 0.1),        # nothing aligned
    (1, 8, 8, 64, 2, 1, 1, 1.0),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv3x3_matches_oracle(P, case):
    B, H, W, Cin, Cout, stride, dil, alpha = case
    x = _rand((B, H, W, Cin), 1)
    k = _rand((3, 3, Cin, Cout), 2, scale=1.0 / np.sqrt(9 * Cin))
    b = _rand((Cout,), 3, scale=0.1)
    y = O.conv2d_same(torch.from_numpy(x), torch.from_numpy(k), torch.from_numpy(b), stride, dil)
    ref = (O.leaky_relu(y, alpha) if alpha != 1.0 else y).numpy()
    out = P.ops.conv3x3(_cuda(x), _cuda(k), _cuda(b), stride=stride, dilation=dil, alpha=alpha).cpu().numpy()
    assert out.shape == ref.shape
    np.testing.assert_allclose(out, ref, atol=2e-5, rtol=1e-5)


def test_conv3x3_strided_views_and_residual(P):
    B, H, W, Cin, Cout = 2, 10, 12, 32, 2
    x = _rand((B, H, W, Cin), 1); k = _rand((3, 3, Cin, Cout), 2, 0.1); b = _rand((Cout,), 3, 0.1)
    res = _rand((B, H, W, 2), 4)
    xin = torch.zeros((B, H, W, 40), device="cuda"); xin[..., 4:36] = _cuda(x)
    buf = torch.zeros((B, H, W, 148), device="cuda"); buf[..., 81:83] = _cuda(res)
    out = torch.zeros((B, H, W, 36), device="cuda")
    P.ops.conv3x3(xin[..., 4:36], _cuda(k), _cuda(b), alpha=1.0, residual=buf[..., 81:83], out=out[..., 32:34])
    ref = O.conv2d_same(torch.from_numpy(x), torch.from_numpy(k), torch.from_numpy(b)).numpy() + res
    np.testing.assert_allclose(out[..., 32:34].cpu().numpy(), ref, atol=2e-5)
    assert (out[..., :32] == 0).all() and (out[..., 34:] == 0).all()


@pytest.mark.parametrize("shape,oh,ow,mul", [((2, 7, 16, 2), 14, 32, 1.0), ((1, 14, 32, 32), 28, 64, 1.0),
                                            ((1, 16, 32, 2), 64, 128, 20.0), ((1, 5, 7, 3), 9, 20, 1.0)])
def test_resize_bilinear_legacy_matches_oracle(P, shape, oh, ow, mul):
    x = _rand(shape, 1)
    ref = (O.resize_bilinear_legacy(torch.from_numpy(x), oh, ow) * mul).numpy()
    out = P.ops.resize_bilinear(_cuda(x), oh, ow, mul).cpu().numpy()
    np.testing.assert_allclose(out, ref, atol=1e-6 * max(1.0, mul), rtol=1e-6)


def test_losses_match_oracle(P):
    B, H, W = 2, 64, 128
    gt = _rand((B, H, W, 2), 1, 5.0)
    pyr = [_rand((B, H >> (6 - l), W >> (6 - l), 2), 10 + l, 0.3) for l in range(5)]
    ff = _rand((B, H, W, 2), 3, 5.0)
    w = O.DEFAULT_LOSS_WEIGHTS
    ref_loss = O.multiscale_loss(torch.from_numpy(gt), [torch.from_numpy(p) for p in pyr], w).item()
    ref_epe = O.EPE(torch.from_numpy(gt), torch.from_numpy(ff)).item()
    loss = P.multiscale_loss(_cuda(gt), [_cuda(p) for p in pyr], w).item()
    epe = P.EPE(_cuda(gt), _cuda(ff)).item()
    assert loss == pytest.approx(ref_loss, rel=1e-5)
    assert epe == pytest.approx(ref_epe, rel=1e-5)
    a, b = _rand((B, 8, 16, 2), 7), _rand((B, 8, 16, 2), 8)
    assert P.L2loss(_cuda(a), _cuda(b)).item() == pytest.approx(O.L2loss(torch.from_numpy(a), torch.from_numpy(b)).item(), rel=1e-5)
    assert P.L1loss(_cuda(a), _cuda(b)).item() == pytest.approx(O.L1loss(torch.from_numpy(a), torch.from_numpy(b)).item(), rel=1e-5)


def test_modules_read_like_the_reference(P):
    """The stand-alone module classes keep the reference's names / signatures (modules.py)."""
    W = O.glorot_weights(2)
    params = {k: _cuda(v) for k, v in W.items()}
    im = _rand((1, 64, 128, 3), 0, 0.3) + 0.5
    pyr = P.modules.FeaturePyramidExtractor_custom(6, params=params)(_cuda(im))
    ref = O.pyramid_extractor(torch.from_numpy(im), W)
    assert [tuple(p.shape) for p in pyr] == [tuple(p.shape) for p in ref]
    for a, b in zip(pyr, ref):
        np.testing.assert_allclose(a.cpu().numpy(), b.numpy(), atol=1e-5)
    f0, f1 = pyr[2], pyr[2].flip(2).contiguous()
    cv = P.modules.CostVolumeLayer(4)(f0, f1)
    flows_up = _cuda(_rand((1,) + tuple(f0.shape[1:3]) + (2,), 5, 0.5))
    feats_up = _cuda(_rand((1,) + tuple(f0.shape[1:3]) + (32,), 6, 0.5))
    est = P.modules.OpticalFlowEstimator_custom(name="optflow_2", params=params)
    flows, fu, featu = est(cv, f0, flows_up, feats_up)
    r_flows, r_fu, r_featu = O.flow_estimator(W, "pwcdcnet/optflow_2", cv.cpu(), f0.cpu(), flows_up.cpu(), feats_up.cpu())
    np.testing.assert_allclose(flows.cpu().numpy(), r_flows.numpy(), atol=1e-5)
    np.testing.assert_allclose(fu.cpu().numpy(), r_fu.numpy(), atol=1e-5)
    np.testing.assert_allclose(featu.cpu().numpy(), r_featu.numpy(), atol=1e-5)
    feats = _cuda(_rand((1, 16, 32, 32), 8, 0.5)); fl = _cuda(_rand((1, 16, 32, 2), 9, 0.5))
    out = P.modules.ContextNetwork(params=params)(fl, feats)
    np.testing.assert_allclose(out.cpu().numpy(), O.context_network(W, fl.cpu(), feats.cpu()).numpy(), atol=1e-5)
    with pytest.raises(AssertionError):
        P.modules.WarpingLayer("cubic")(f0, flows_up)


def test_host_layer_rejects_bad_inputs(P):
    with pytest.raises(TypeError):
        P.ops.cost_volume(torch.zeros(1, 8, 8, 32, device="cuda", dtype=torch.float16), torch.zeros(1, 8, 8, 32, device="cuda"))
    with pytest.raises(ValueError):
        P.ops.cost_volume(torch.zeros(1, 8, 8, 32, device="cuda"), torch.zeros(1, 8, 9, 32, device="cuda"))
    with pytest.raises(ValueError):   # NCHW-permuted view is not pixel-strided NHWC
        P.ops.cost_volume(torch.zeros(1, 32, 8, 8, device="cuda").permute(0, 2, 3, 1), torch.zeros(1, 8, 8, 32, device="cuda"))
    with pytest.raises(P.ops.PwcError):
        P.ops.cost_volume(torch.zeros(1, 8, 8, 32), torch.zeros(1, 8, 8, 32))
    with pytest.raises(P.ops.PwcError):   # C not a multiple of 4 -> PWC_E_ALIGN from the C entry point
        P.ops.cost_volume(torch.zeros(1, 8, 8, 30, device="cuda"), torch.zeros(1, 8, 8, 30, device="cuda"))


TC_CASES = [
    # (B, H, W, Cin, channel stride, Cout, dilation[, stride])
    (1, 8, 16, 32, 32, 32, 1),
    (2, 32, 48, 16, 16, 16, 1),          # Cin = 16: 64-byte swizzle path (pyramid level 1)
    (2, 32, 48, 16, 16, 32, 1, 2),       # stride 2 through TMA element strides, asymmetric SAME (0 before, 1 after)
    (1, 28, 64, 64, 64, 96, 1, 2),
    (1, 15, 17, 32, 32, 64, 1, 2),       # odd sizes: SAME pads 1 before / 1 after
    (2, 14, 32, 128, 128, 192, 1, 2),
    (2, 14, 32, 128, 128, 128, 1),
    (1, 28, 64, 147, 148, 128, 1),      # estimator conv 0 at level 4: ragged K (147 -> zero-filled to 160)
    (1, 7, 16, 273, 276, 128, 1),       # coarsest level: 7 rows in an 8-row tile
    (1, 28, 64, 128, 128, 96, 2),
    (1, 24, 40, 96, 96, 64, 16),        # dilation 16: taps mostly outside the image (TMA zero fill)
    (2, 9, 21, 64, 64, 32, 4),          # ragged tiles in x and y
    (1, 16, 32, 192, 192, 192, 1),      # N = 192 (TMEM allocation 256 columns)
]


@pytest.mark.parametrize("n_split,tol", [(3, 3e-5), (1, 1e-2)])
@pytest.mark.parametrize("case", TC_CASES)
def test_conv3x3_tcgen05_matches_oracle(P, case, n_split, tol):
    """tcgen05 implicit-GEMM conv vs the oracle's fp32 conv.  3xTF32 must be fp32-class (3e-5 max-abs on
    O(1) outputs); plain TF32 (operands truncated to 10 mantissa bits by the hardware) within 1e-2."""
    from pwcnet_b200 import ops_tc
    B, H, W, Cin, cs, Cout, dil = case[:7]
    stride = case[7] if len(case) > 7 else 1
    buf = _rand((B, H, W, cs), 1)
    k = _rand((3, 3, Cin, Cout), 2, scale=1.0 / np.sqrt(9 * Cin))
    b = _rand((Cout,), 3, scale=0.1)
    x = buf[..., :Cin]
    ref = O.leaky_relu(O.conv2d_same(torch.from_numpy(np.ascontiguousarray(x)), torch.from_numpy(k), torch.from_numpy(b), stride, dil), 0.1).numpy()
    wp = ops_tc.pack_weights(_cuda(k))
    out = ops_tc.conv3x3_tc(_cuda(buf)[..., :Cin], wp, _cuda(b), Cin, Cout, dilation=dil, alpha=0.1, n_split=n_split,
                            stride=stride)
    assert out.shape == ref.shape
    np.testing.assert_allclose(out.cpu().numpy(), ref, atol=tol, rtol=0)


def test_conv3x3_tcgen05_writes_only_its_slot(P):
    from pwcnet_b200 import ops_tc
    B, H, W, Cin, Cout = 1, 12, 20, 64, 32
    x = _rand((B, H, W, Cin), 1); k = _rand((3, 3, Cin, Cout), 2, 0.05); b = _rand((Cout,), 3, 0.1)
    out = torch.full((B, H, W, 40), 9.0, device="cuda")
    ops_tc.conv3x3_tc(_cuda(x), ops_tc.pack_weights(_cuda(k)), _cuda(b), Cin, Cout, alpha=1.0, n_split=3, out=out[..., 4:36])
    ref = O.conv2d_same(torch.from_numpy(x), torch.from_numpy(k), torch.from_numpy(b)).numpy()
    np.testing.assert_allclose(out[..., 4:36].cpu().numpy(), ref, atol=2e-5)
    assert (out[..., :4] == 9.0).all() and (out[..., 36:] == 9.0).all()


@pytest.mark.parametrize("case", TC_CASES + [(1, 28, 64, 34, 36, 128, 1), (2, 14, 32, 128, 128, 128, 1)])
def test_conv3x3_tcgen05_f16x3_matches_oracle(P, case):
    """tcgen05 kind::f16 conv with the 3 x fp16 scaled-residual split (the default path): fp32-class,
    max-abs 2e-5 on O(1) outputs vs the oracle's fp32 conv."""
    from pwcnet_b200 import ops_tc
    B, H, W, Cin, cs, Cout, dil = case[:7]
    stride = case[7] if len(case) > 7 else 1
    buf = _rand((B, H, W, cs), 1)
    k = _rand((3, 3, Cin, Cout), 2, scale=1.0 / np.sqrt(9 * Cin))
    b = _rand((Cout,), 3, scale=0.1)
    x = buf[..., :Cin]
    ref = O.leaky_relu(O.conv2d_same(torch.from_numpy(np.ascontiguousarray(x)), torch.from_numpy(k), torch.from_numpy(b), stride, dil), 0.1).numpy()
    out = ops_tc.conv3x3_tc_f16(_cuda(buf)[..., :Cin], ops_tc.pack_weights_f16(_cuda(k)), _cuda(b), Cin, Cout, dilation=dil,
                                alpha=0.1, stride=stride)
    assert out.shape == ref.shape
    np.testing.assert_allclose(out.cpu().numpy(), ref, atol=2e-5, rtol=0)


@pytest.mark.parametrize("case", [(2, 14, 32, 128, 128, 192, 1, 2), (1, 16, 32, 192, 192, 192, 1), (1, 14, 32, 243, 244, 128, 1),
                                  (1, 28, 64, 64, 64, 96, 1, 2), (2, 7, 16, 192, 192, 192, 1), (1, 24, 40, 96, 96, 64, 16)])
def test_conv3x3_f16x3_split_k_cluster_matches_oracle(P, case, monkeypatch):
    """Opt-in split-K of the streaming kernel (PWC_TC_KSPLIT=1: clusters of 3 CTAs, one kernel row of taps each, partial
    tiles reduced through rank 0's shared memory) on sub-wave grids: same fp32-class bar as the default path, and it
    writes only its channel slot."""
    from pwcnet_b200 import ops_tc
    B, H, W, Cin, cs, Cout, dil = case[:7]
    stride = case[7] if len(case) > 7 else 1
    buf = _rand((B, H, W, cs), 1)
    k = _rand((3, 3, Cin, Cout), 2, scale=1.0 / np.sqrt(9 * Cin))
    b = _rand((Cout,), 3, scale=0.1)
    x = buf[..., :Cin]
    ref = O.leaky_relu(O.conv2d_same(torch.from_numpy(np.ascontiguousarray(x)), torch.from_numpy(k), torch.from_numpy(b), stride, dil), 0.1).numpy()
    wp = ops_tc.pack_weights_f16(_cuda(k))
    monkeypatch.setenv("PWC_CONV_HALO", "0")     # the streaming kernel, also for the stride-1 shapes
    base = ops_tc.conv3x3_tc_f16(_cuda(buf)[..., :Cin], wp, _cuda(b), Cin, Cout, dilation=dil, alpha=0.1, stride=stride)
    monkeypatch.setenv("PWC_TC_KSPLIT", "1")
    oh, ow = ref.shape[1:3]
    out = torch.full((B, oh, ow, Cout + 8), 7.0, device="cuda")
    ops_tc.conv3x3_tc_f16(_cuda(buf)[..., :Cin], wp, _cuda(b), Cin, Cout, dilation=dil, alpha=0.1, stride=stride, out=out[..., 4:4 + Cout])
    np.testing.assert_allclose(out[..., 4:4 + Cout].cpu().numpy(), ref, atol=2e-5, rtol=0)
    np.testing.assert_allclose(out[..., 4:4 + Cout].cpu().numpy(), base.cpu().numpy(), atol=1e-5, rtol=0)
    assert (out[..., :4] == 7.0).all() and (out[..., 4 + Cout:] == 7.0).all()


HALO_CASES = [  # rows of >= 96 pixels, stride 1, Cout <= 128 take the halo-resident kernel (conv_tc_halo.cu)
    # (B, H, W, Cin, channel stride, Cout, dilation)
    (1, 4, 128, 32, 32, 32, 1), (1, 9, 130, 32, 32, 16, 1), (2, 14, 256, 128, 128, 128, 1), (1, 7, 200, 147, 148, 128, 1),
    (1, 5, 96, 64, 64, 64, 1), (2, 6, 300, 16, 16, 16, 1), (1, 3, 512, 96, 96, 48, 1), (1, 2, 128, 34, 36, 128, 1),
    (1, 9, 130, 32, 32, 32, 2), (2, 20, 256, 128, 128, 128, 4), (1, 30, 200, 128, 128, 96, 8), (1, 40, 256, 96, 96, 64, 16),
    (1, 5, 100, 64, 64, 32, 16), (3, 1, 97, 32, 32, 16, 1), (1, 12, 160, 16, 16, 32, 2), (1, 10, 140, 16, 20, 16, 4),   # 64-byte-row tiles, dilated
    # rows narrower than a tile: flat mode packs several rows into the 128 MMA rows
    (2, 7, 16, 32, 32, 32, 1), (1, 14, 32, 243, 244, 128, 1), (2, 28, 64, 96, 96, 64, 1), (1, 9, 21, 64, 64, 32, 1),
    (3, 5, 8, 32, 32, 16, 1), (1, 30, 100, 48, 48, 96, 1), (2, 13, 45, 16, 16, 16, 1), (1, 1, 9, 32, 32, 32, 1),
]


@pytest.mark.parametrize("case", HALO_CASES)
def test_conv3x3_halo_resident_kernel_matches_oracle_and_streaming_kernel(P, case, monkeypatch):
    """Halo-resident tcgen05 conv: one 3 x (128+2d) box per channel slice, shifted A descriptors, persistent CTA.
    Must agree with the oracle's fp32 conv (2e-5 max-abs on O(1) outputs) and with the streaming kernel, and write
    only its channel slot."""
    from pwcnet_b200 import ops_tc
    B, H, W, Cin, cs, Cout, dil = case
    buf = _rand((B, H, W, cs), 1)
    k = _rand((3, 3, Cin, Cout), 2, scale=1.0 / np.sqrt(9 * Cin))
    b = _rand((Cout,), 3, scale=0.1)
    x = buf[..., :Cin]
    ref = O.leaky_relu(O.conv2d_same(torch.from_numpy(np.ascontiguousarray(x)), torch.from_numpy(k), torch.from_numpy(b), 1, dil), 0.1).numpy()
    wp = ops_tc.pack_weights_f16(_cuda(k))
    monkeypatch.setenv("PWC_CONV_HALO", "1")
    out = torch.full((B, H, W, Cout + 8), 9.0, device="cuda")
    ops_tc.conv3x3_tc_f16(_cuda(buf)[..., :Cin], wp, _cuda(b), Cin, Cout, dilation=dil, alpha=0.1, out=out[..., 4:4 + Cout])
    np.testing.assert_allclose(out[..., 4:4 + Cout].cpu().numpy(), ref, atol=2e-5, rtol=0)
    assert (out[..., :4] == 9.0).all() and (out[..., 4 + Cout:] == 9.0).all()
    monkeypatch.setenv("PWC_CONV_HALO", "0")
    stream = ops_tc.conv3x3_tc_f16(_cuda(buf)[..., :Cin], wp, _cuda(b), Cin, Cout, dilation=dil, alpha=0.1)
    np.testing.assert_allclose(out[..., 4:4 + Cout].cpu().numpy(), stream.cpu().numpy(), atol=2e-5, rtol=0)


@pytest.mark.parametrize("shape", [(2, 7, 16, 32, 32), (1, 28, 64, 32, 36), (1, 20, 256, 32, 36), (1, 5, 130, 64, 64)])
def test_flow_head_on_tensor_cores_matches_oracle(P, shape):
    """2-channel flow heads (modules.py:274-277, 325-326): kernel zero-padded to 16 output channels, only 2 stored,
    residual added after the (absent) activation, written into a 2-channel slot of a wider buffer."""
    from pwcnet_b200 import ops_tc
    B, H, W, Cin, cs = shape
    buf = _rand((B, H, W, cs), 1); x = buf[..., :Cin]
    k = _rand((3, 3, Cin, 2), 2, 1.0 / np.sqrt(9 * Cin)); b = _rand((2,), 3, 0.1); res = _rand((B, H, W, 2), 4)
    ref = (O.conv2d_same(torch.from_numpy(np.ascontiguousarray(x)), torch.from_numpy(k), torch.from_numpy(b)) + torch.from_numpy(res)).numpy()
    kp = torch.zeros((3, 3, Cin, 16), device="cuda"); kp[..., :2] = _cuda(k)
    bp = torch.zeros(16, device="cuda"); bp[:2] = _cuda(b)
    out = torch.full((B, H, W, 6), 9.0, device="cuda")
    ops_tc.conv3x3_tc_f16_head(_cuda(buf)[..., :Cin], ops_tc.pack_weights_f16(kp), bp, Cin, 2, 16, alpha=1.0,
                               residual=_cuda(res), out=out[..., 2:4])
    np.testing.assert_allclose(out[..., 2:4].cpu().numpy(), ref, atol=2e-5, rtol=0)
    assert (out[..., :2] == 9.0).all() and (out[..., 4:] == 9.0).all()


def test_conv3x3_f16x3_small_and_large_magnitudes(P):
    """The scaled residual keeps accuracy relative to the data scale from 1e-3 to 1e2."""
    from pwcnet_b200 import ops_tc
    for scale in (1e-3, 1.0, 1e2):
        x = _rand((1, 16, 32, 64), 1, scale); k = _rand((3, 3, 64, 64), 2, 1.0 / 24); b = np.zeros(64, np.float32)
        ref = O.conv2d_same(torch.from_numpy(x), torch.from_numpy(k), torch.from_numpy(b)).numpy()
        out = ops_tc.conv3x3_tc_f16(_cuda(x), ops_tc.pack_weights_f16(_cuda(k)), _cuda(b), 64, 64, alpha=1.0)
        np.testing.assert_allclose(out.cpu().numpy(), ref, atol=2e-5 * scale, rtol=0)


@pytest.mark.parametrize("shape", [(2, 64, 128), (1, 70, 132), (3, 18, 260), (1, 448, 1024)])
@pytest.mark.parametrize("u8", [False, True])
def test_conv_first_matches_oracle(P, shape, u8):
    """First pyramid conv (3 -> 16, stride 2, SAME, leaky 0.1; modules.py:62-63) from float32 or uint8 images: exact-fp32 class
    vs the oracle's conv, ragged tiles, strided destination; the uint8 path equals the float path on the /255.0 feed bit for bit."""
    B, H, W = shape
    rng = np.random.default_rng(H + W)
    k = rng.uniform(-0.3, 0.3, (3, 3, 3, 16)).astype(np.float32)
    b = rng.normal(0, 0.1, 16).astype(np.float32)
    xu = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    xf = (xu / 255.0).astype(np.float32)
    ref = O.leaky_relu(O.conv2d_same(torch.from_numpy(xf), torch.from_numpy(k), torch.from_numpy(b), stride=2), 0.1).numpy()
    buf = torch.full((B, H // 2, W // 2, 20), 7.0, device="cuda")
    x = torch.from_numpy(xu if u8 else xf).cuda()
    out = P.ops.conv_first(x, _cuda(k), _cuda(b), 0.1, out=buf[..., 2:18])
    np.testing.assert_allclose(out.cpu().numpy(), ref, atol=2e-6, rtol=1e-6)
    assert float((buf[..., :2] - 7.0).abs().max()) == 0 and float((buf[..., 18:] - 7.0).abs().max()) == 0
    dense = P.ops.conv_first(torch.from_numpy(xf).cuda(), _cuda(k), _cuda(b), 0.1)
    assert torch.equal(dense, out.contiguous())
    with pytest.raises(ValueError):
        P.ops.conv_first(torch.zeros((1, 63, 128, 3), device="cuda"), _cuda(k), _cuda(b))


@pytest.mark.parametrize("shape,cmid,cout,dil", [((2, 20, 140, 36), 128, 96, 1), ((1, 56, 128, 128), 64, 32, 4), ((1, 9, 17, 64), 32, 32, 1),
                                                  ((1, 40, 160, 96), 96, 64, 16)])
def test_conv_chain_with_split_activations_is_bit_identical(P, shape, cmid, cout, dil):
    """conv -> conv with the intermediate tensor stored as [h | l] fp16 rows (pwc_conv3x3_tc_f16_split_fwd): the split tensor
    reproduces the fp32 intermediate to fp32 precision and the chain's result equals the fp32-tensor chain bit for bit."""
    from pwcnet_b200 import ops_tc
    B, H, W, C = shape
    x = _cuda(_rand(shape, 1))
    k1, b1 = _cuda(_rand((3, 3, C, cmid), 2, 0.05)), _cuda(_rand((cmid,), 3, 0.1))
    k2, b2 = _cuda(_rand((3, 3, cmid, cout), 4, 0.05)), _cuda(_rand((cout,), 5, 0.1))
    w1, w2 = ops_tc.pack_weights_f16(k1), ops_tc.pack_weights_f16(k2)
    mid = ops_tc.conv3x3_tc_f16(x, w1, b1, C, cmid, dilation=dil, alpha=0.1)
    ref = ops_tc.conv3x3_tc_f16(mid, w2, b2, cmid, cout, dilation=dil, alpha=0.1)
    mid_s = torch.empty((B, H, W, 2 * cmid), dtype=torch.float16, device="cuda")
    mid_f = torch.empty((B, H, W, cmid), device="cuda")
    ops_tc.conv3x3_tc_f16_split(x, w1, b1, C, cmid, dilation=dil, alpha=0.1, out=mid_f, out_split=mid_s)
    assert torch.equal(mid_f, mid)
    hl = mid_s.float().view(B, H, W, cmid // 32, 2, 32)
    rec = (hl[..., 0, :] + hl[..., 1, :] / 2048.0).reshape(B, H, W, cmid)
    np.testing.assert_allclose(rec.cpu().numpy(), mid.cpu().numpy(), rtol=2e-7, atol=1e-9)
    out = ops_tc.conv3x3_tc_f16_split(mid_s, w2, b2, cmid, cout, dilation=dil, alpha=0.1)
    assert torch.equal(out, ref)
    with pytest.raises(ValueError):
        ops_tc.conv3x3_tc_f16_split(mid_s[..., :cmid], w2, b2, cmid, cout)


@pytest.mark.parametrize("shape,cout", [((2, 40, 264, 16), 32), ((1, 56, 128, 32), 64), ((2, 28, 64, 64), 96), ((3, 14, 34, 96), 128),
                                        ((1, 6, 8, 16), 16), ((1, 300, 2, 32), 32)])
def test_conv_stride2_space_to_depth_matches_oracle(P, shape, cout):
    """Stride-2 conv + bias + leaky (modules.py:62-63) on the halo kernel as a 2x2 conv over the space-to-depth view
    (pwc_conv3x3_s2d_reindex + pwc_conv3x3_s2_tc_f16_fwd): vs the oracle's TF-'SAME' stride-2 conv, fp32-class tolerance;
    the split-row output reproduces the fp32 output.  Covers row tiles with a ragged right edge, the flat mode (rows of
    fewer than 128 outputs), one-column outputs, every pyramid channel pair."""
    from pwcnet_b200 import ops_tc
    B, H, W, C = shape
    x, k, b = _rand(shape, 11), _rand((3, 3, C, cout), 12, 0.05), _rand((cout,), 13, 0.1)
    ref = O.leaky_relu(O.conv2d_same(torch.from_numpy(x), torch.from_numpy(k), torch.from_numpy(b), stride=2), 0.1).numpy()
    k2 = ops_tc.s2d_reindex(_cuda(k))
    # the re-indexed kernel: taps (0..1, 0..1) hold W[2dy+py][2dx+px], everything else is zero
    k2n = k2.cpu().numpy().reshape(3, 3, 2, 2, C, cout)
    for dy in range(3):
        for dx in range(3):
            for py in range(2):
                for px in range(2):
                    ky, kx = 2 * dy + py, 2 * dx + px
                    want = k[ky, kx] if (dy < 2 and dx < 2 and ky < 3 and kx < 3) else np.zeros((C, cout), np.float32)
                    assert np.array_equal(k2n[dy, dx, py, px], want)
    wp = ops_tc.pack_weights_f16(k2)
    y = ops_tc.conv3x3_s2_tc_f16(_cuda(x), wp, _cuda(b), C, cout, alpha=0.1)
    assert tuple(y.shape) == (B, H // 2, W // 2, cout)
    np.testing.assert_allclose(y.cpu().numpy(), ref, rtol=0, atol=2e-5 * max(1.0, float(np.abs(ref).max())))
    # into a channel slot of a wider buffer, and as split rows
    wide = torch.full((B, H // 2, W // 2, cout + 16), 7.0, device="cuda")
    ops_tc.conv3x3_s2_tc_f16(_cuda(x), wp, _cuda(b), C, cout, alpha=0.1, out=wide[..., 16:])
    assert torch.equal(wide[..., 16:], y) and bool((wide[..., :16] == 7.0).all())
    if cout % 32 == 0:
        ys = torch.empty((B, H // 2, W // 2, 2 * cout), dtype=torch.float16, device="cuda")
        ops_tc.conv3x3_s2_tc_f16(_cuda(x), wp, _cuda(b), C, cout, alpha=0.1, out_split=ys)
        hl = ys.float().view(B, H // 2, W // 2, cout // 32, 2, 32)
        rec = (hl[..., 0, :] + hl[..., 1, :] / 2048.0).reshape(B, H // 2, W // 2, cout)
        np.testing.assert_allclose(rec.cpu().numpy(), y.cpu().numpy(), rtol=2e-7, atol=1e-9)
    with pytest.raises(ValueError):
        ops_tc.conv3x3_s2_tc_f16(_cuda(x)[:, :-1], wp, _cuda(b), C, cout)          # odd height
