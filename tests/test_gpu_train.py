"""GPU parity of the backward kernels and of the whole training step (reference train.py:65-92) against the
CPU oracle.  The oracle restates the reference forward in torch-CPU, so torch autograd over it gives exactly the
gradients TF-1.8's autodiff produced for the reference graph (same op decomposition); Adam is
checked against the oracle's restatement of tf.train.AdamOptimizer.

Tolerances: fp32 kernels that differ from the oracle in summation order (and atomics order) only:
max-abs error <= 2e-5 * max|reference| + 1e-6 per tensor for single operators, 2e-4 relative for gradients
that went through the whole network."""
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
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _close(got, ref, rel=2e-5, name=""):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    ref = ref.detach().cpu().numpy() if isinstance(ref, torch.Tensor) else np.asarray(ref)
    assert got.shape == ref.shape, f"{name}: shape {got.shape} vs {ref.shape}"
    tol = rel * max(float(np.abs(ref).max()), 1e-30) + 1e-6
    err = float(np.abs(got - ref).max())
    assert err <= tol, f"{name}: max-abs err {err:.3e} > tol {tol:.3e} (max|ref| {np.abs(ref).max():.3e})"


CONV_CASES = [  # B, H, W, Cin, Cout, stride, dilation
    (2, 12, 20, 3, 16, 2, 1), (1, 16, 24, 16, 32, 2, 1), (2, 9, 13, 32, 32, 1, 1), (1, 11, 18, 34, 128, 1, 2),
    (1, 20, 40, 96, 64, 1, 8), (1, 7, 16, 128, 96, 1, 16), (2, 8, 16, 32, 2, 1, 1), (1, 10, 12, 147, 128, 1, 1),
    (1, 13, 17, 64, 70, 2, 1), (1, 1, 2, 192, 192, 1, 1), (1, 17, 45, 16, 16, 1, 1), (1, 8, 70, 20, 12, 1, 1), (2, 9, 33, 4, 2, 2, 1),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_dgrad_wgrad_match_autograd(P, case):
    from pwcnet_b200 import ops_bwd
    B, H, W, Cin, Cout, s, d = case
    x = torch.from_numpy(_rand((B, H, W, Cin), 1)).requires_grad_(True)
    k = torch.from_numpy(_rand((3, 3, Cin, Cout), 2, 0.1)).requires_grad_(True)
    b = torch.from_numpy(_rand((Cout,), 3, 0.1)).requires_grad_(True)
    y = O.conv2d_same(x, k, b, s, d)
    dy = torch.from_numpy(_rand(tuple(y.shape), 4))
    y.backward(dy)
    dx = torch.full((B, H, W, Cin), 7.0, device="cuda")          # overwritten when accumulate=False
    ops_bwd.conv3x3_dgrad(_cuda(dy.numpy()), _cuda(k.detach().numpy()), dx, stride=s, dilation=d)
    _close(dx, x.grad, name="dx")
    dw = torch.zeros((3, 3, Cin, Cout), device="cuda")
    db = torch.zeros((Cout,), device="cuda")
    ops_bwd.conv3x3_wgrad(_cuda(x.detach().numpy()), _cuda(dy.numpy()), dw, db, stride=s, dilation=d)
    _close(dw, k.grad, name="dw")
    _close(db, b.grad, name="db")


WGRAD_TC_CASES = [  # B, H, W, Cin, Cout, stride, dilation
    (2, 9, 13, 32, 64, 1, 1), (1, 11, 18, 36, 128, 1, 2), (1, 20, 40, 96, 64, 1, 8), (1, 7, 16, 128, 96, 1, 16),
    (1, 10, 12, 148, 128, 1, 1), (1, 13, 17, 64, 96, 2, 1), (1, 3, 5, 192, 192, 1, 1), (2, 16, 70, 64, 32, 1, 1),
    (1, 12, 300, 128, 128, 1, 1), (2, 14, 32, 276, 128, 1, 1), (1, 16, 24, 16, 32, 2, 1), (2, 20, 36, 16, 16, 1, 1),
    (1, 9, 40, 20, 48, 1, 1),
]


@pytest.mark.parametrize("case", WGRAD_TC_CASES)
def test_conv_wgrad_tensor_core_matches_autograd(P, case):
    """tsplit (transpose + fp16 split, bias gradient) + tcgen05 wgrad vs torch autograd of the oracle conv; fp32-class:
    2e-5 relative to the largest gradient entry."""
    from pwcnet_b200 import ops_bwd
    B, H, W, Cin, Cout, s, d = case
    x = torch.from_numpy(_rand((B, H, W, Cin), 1))
    k = torch.from_numpy(_rand((3, 3, Cin, Cout), 2, 0.1)).requires_grad_(True)
    b = torch.from_numpy(_rand((Cout,), 3, 0.1)).requires_grad_(True)
    y = O.conv2d_same(x, k, b, s, d)
    dy = torch.from_numpy(_rand(tuple(y.shape), 4, 1e-3))          # gradients are small: exercises fp16 subnormals of h
    y.backward(dy)
    xb = torch.zeros((B, H, W, Cin + 4), device="cuda"); xb[..., :Cin] = x.cuda()      # a slot view, like the concat buffers
    db = torch.zeros((Cout,), device="cuda")
    xT = ops_bwd.tsplit(xb[..., :Cin], conv_input=True, stride=s, dilation=d)
    dyT = ops_bwd.tsplit(dy.cuda(), db=db)
    dw = torch.zeros((3, 3, Cin, Cout), device="cuda")
    ops_bwd.conv3x3_wgrad_tc(xT, dyT, dw, (B, H, W, Cin), Cout, stride=s, dilation=d)
    _close(dw, k.grad, name="dw (tensor cores)")
    _close(db, b.grad, name="db (tsplit)")


def test_conv_dgrad_mask_accumulate_and_slots(P):
    """Leaky mask in the epilogue, accumulation, and reading/writing channel slots of wider buffers."""
    from pwcnet_b200 import ops_bwd
    B, H, W, Cin, Cout = 2, 10, 14, 36, 32
    x = torch.from_numpy(_rand((B, H, W, Cin), 1))
    k = torch.from_numpy(_rand((3, 3, Cin, Cout), 2, 0.1))
    dy = torch.from_numpy(_rand((B, H, W, Cout), 4))
    prev = torch.from_numpy(_rand((B, H, W, Cin), 5))
    xr = x.clone().requires_grad_(True)
    # x is the (post-activation) output of a leaky layer: y_prev = leaky(z); d/dz = d/dx * leaky'(x)
    y = O.conv2d_same(xr, k, None, 1, 1)
    y.backward(dy)
    ref = prev + xr.grad * torch.where(x > 0, torch.ones_like(x), torch.full_like(x, 0.1))
    dybuf = torch.zeros((B, H, W, 40), device="cuda"); dybuf[..., 4:36] = _cuda(dy.numpy())
    dxbuf = torch.zeros((B, H, W, 48), device="cuda"); dxbuf[..., 8:44] = _cuda(prev.numpy())
    ops_bwd.conv3x3_dgrad(dybuf[..., 4:36], _cuda(k.numpy()), dxbuf[..., 8:44], mask=_cuda(x.numpy()), mask_alpha=0.1,
                          accumulate=True)
    _close(dxbuf[..., 8:44], ref, name="dx slot")
    assert float(dxbuf[..., :8].abs().max()) == 0 and float(dxbuf[..., 44:].abs().max()) == 0


@pytest.mark.parametrize("use_mask", [True, False])
@pytest.mark.parametrize("shape,cout,accumulate", [((1, 6, 160, 64), 32, True), ((2, 5, 300, 128), 128, False), ((1, 4, 130, 32), 64, True),
                                                   ((1, 3, 129, 16), 16, False), ((1, 5, 200, 16), 32, True), ((2, 3, 140, 32), 32, False)])
def test_conv_dgrad_wide_rows_tma_epilogue(P, shape, cout, accumulate, use_mask):
    """Row-tile mode of the halo kernel (rows of >= 128 pixels: the level-2/3 layers of a training step): the leaky mask tile
    arrives through TMA and `accumulate` is a reduce-add bulk store (conv_tc_halo.cu, hl_epilogue_pass); ragged right edges;
    without a mask the 16- and 32-channel cases take the software-pipelined epilogue unless they accumulate."""
    from pwcnet_b200 import ops_bwd
    B, H, W, Cin = shape
    x = torch.from_numpy(_rand((B, H, W, Cin), 1))
    k = torch.from_numpy(_rand((3, 3, Cin, cout), 2, 0.1))
    dy = torch.from_numpy(_rand((B, H, W, cout), 4))
    prev = torch.from_numpy(_rand((B, H, W, Cin), 5))
    xr = x.clone().requires_grad_(True)
    O.conv2d_same(xr, k, None, 1, 1).backward(dy)
    ref = xr.grad * torch.where(x > 0, torch.ones_like(x), torch.full_like(x, 0.1)) if use_mask else xr.grad
    if accumulate:
        ref = ref + prev
    dxbuf = torch.zeros((B, H, W, Cin + 16), device="cuda"); dxbuf[..., 8:8 + Cin] = _cuda(prev.numpy())
    ops_bwd.conv3x3_dgrad(_cuda(dy.numpy()), _cuda(k.numpy()), dxbuf[..., 8:8 + Cin], mask=_cuda(x.numpy()) if use_mask else None,
                          mask_alpha=0.1, accumulate=accumulate)
    _close(dxbuf[..., 8:8 + Cin], ref, name="dx (row tiles)")
    assert float(dxbuf[..., :8].abs().max()) == 0 and float(dxbuf[..., 8 + Cin:].abs().max()) == 0


def test_conv_wgrad_cin_map(P):
    """Concat layers: x in internal channel order with padding channels, dw in the reference order."""
    from pwcnet_b200 import ops_bwd
    B, H, W, Cref, Cout = 1, 9, 11, 10, 16
    perm = [3, 4, -1, 0, 1, 2, 9, 8, 7, 6, 5, -1]              # internal channel i holds reference channel perm[i]
    xr = torch.from_numpy(_rand((B, H, W, Cref), 1)).requires_grad_(False)
    k = torch.from_numpy(_rand((3, 3, Cref, Cout), 2, 0.1)).requires_grad_(True)
    y = O.conv2d_same(xr, k, None, 1, 1)
    dy = torch.from_numpy(_rand(tuple(y.shape), 4))
    y.backward(dy)
    xi = torch.zeros((B, H, W, len(perm)))
    for i, r in enumerate(perm):
        if r >= 0:
            xi[..., i] = xr[..., r]
    dw = torch.zeros((3, 3, Cref, Cout), device="cuda")
    ops_bwd.conv3x3_wgrad(xi.cuda(), _cuda(dy.numpy()), dw, None, cin_map=torch.tensor(perm, dtype=torch.int32, device="cuda"))
    _close(dw, k.grad, name="dw (cin_map)")
    kint = torch.zeros((3, 3, len(perm), Cout), device="cuda")
    ops_bwd.permute_cin(_cuda(k.detach().numpy()), kint, torch.tensor(perm, dtype=torch.int32, device="cuda"))
    for i, r in enumerate(perm):
        ref = k.detach()[:, :, r, :] if r >= 0 else torch.zeros((3, 3, Cout))
        assert torch.equal(kint[:, :, i, :].cpu(), ref)


def test_rot_weights_turns_forward_conv_into_dgrad(P):
    from pwcnet_b200 import ops_bwd
    B, H, W, Cin, Cout, d = 1, 12, 16, 32, 48, 2
    k = _rand((3, 3, Cin, Cout), 2, 0.1)
    dy = _rand((B, H, W, Cout), 4)
    dx = torch.empty((B, H, W, Cin), device="cuda")
    ops_bwd.conv3x3_dgrad(_cuda(dy), _cuda(k), dx, dilation=d)
    krot = ops_bwd.rot_weights(_cuda(k))
    assert krot.shape == (3, 3, Cout, Cin)
    via_fwd = P.ops.conv3x3(_cuda(dy), krot, torch.zeros(Cin, device="cuda"), dilation=d, alpha=1.0)
    _close(via_fwd, dx, name="rot")
    part = ops_bwd.rot_weights(_cuda(k), ci_begin=8, ci_count=20, ci_pad=32)
    assert torch.equal(part[..., :20], krot[..., 8:28]) and float(part[..., 20:].abs().max()) == 0


@pytest.mark.parametrize("case", [(2, 9, 13, 32, 32, 1), (1, 11, 18, 36, 128, 2), (1, 20, 40, 96, 64, 8), (1, 10, 12, 148, 128, 1),
                                  (1, 7, 16, 276, 128, 1), (2, 16, 16, 16, 16, 1), (1, 12, 20, 128, 96, 16)])
def test_tc_dgrad_matches_cuda_core_dgrad(P, case):
    """Stride-1 dgrad on tcgen05 (rotated kernel, 3 x fp16 split) vs the exact fp32 dgrad kernel: channel counts
    that are not multiples of 16, wider than 256 (split in parts), leaky mask, accumulation, slot views."""
    from pwcnet_b200 import ops_bwd, ops_tc
    B, H, W, Cdx, Cdy, d = case
    k = _cuda(_rand((3, 3, Cdx, Cdy), 2, 0.1))
    dy = _cuda(_rand((B, H, W, Cdy), 4))
    mask = _cuda(_rand((B, H, W, Cdx), 5))
    prev = _rand((B, H, W, Cdx), 6)
    ref = _cuda(prev)
    ops_bwd.conv3x3_dgrad(dy, k, ref, dilation=d, mask=mask, accumulate=True)
    buf = torch.zeros((B, H, W, Cdx + 8), device="cuda")
    got = buf[..., 4:4 + Cdx]
    got.copy_(_cuda(prev))
    n_parts = -(-Cdx // 256)
    size = (-(-Cdx // n_parts) + 15) // 16 * 16
    for c0 in range(0, Cdx, size):
        cnt = min(size, Cdx - c0)
        pad = (cnt + 15) // 16 * 16
        packed = ops_tc.pack_weights_f16(ops_bwd.rot_weights(k, ci_begin=c0, ci_count=cnt, ci_pad=pad))
        ops_bwd.conv3x3_tc_f16_dgrad(dy, packed, got[..., c0:c0 + cnt], pad, dilation=d, mask=mask[..., c0:c0 + cnt],
                                     accumulate=True)
    _close(got, ref, rel=2e-5, name="tc dgrad")
    assert float(buf[..., :4].abs().max()) == 0 and float(buf[..., 4 + Cdx:].abs().max()) == 0


@pytest.mark.parametrize("case", [(2, 12, 16, 16, 32), (1, 11, 13, 32, 64), (1, 24, 32, 128, 196), (1, 7, 10, 64, 96)])
def test_stride2_dgrad_through_zero_insertion_matches_cuda_core_dgrad(P, case):
    """Stride-2 dgrad = stride-1 tcgen05 dgrad of the zero-inserted dy (even and odd H/W: pad_top 0 / 1), vs the exact
    fp32 parity-class kernel, accumulating into an existing gradient."""
    from pwcnet_b200 import ops_bwd, ops_tc
    B, H, W, Cdx, Cdy = case
    OH, OW = (H + 1) // 2, (W + 1) // 2
    k = _cuda(_rand((3, 3, Cdx, Cdy), 2, 0.1))
    dy = _cuda(_rand((B, OH, OW, Cdy), 4))
    prev = _rand((B, H, W, Cdx), 6)
    ref = _cuda(prev)
    ops_bwd.conv3x3_dgrad(dy, k, ref, stride=2, accumulate=True)
    pad_t, pad_l = max((OH - 1) * 2 + 3 - H, 0) // 2, max((OW - 1) * 2 + 3 - W, 0) // 2
    dil = ops_bwd.dilate2(dy, torch.full((B, H, W, Cdy), 7.0, device="cuda"), 1 - pad_t, 1 - pad_l)
    assert int((dil != 0).sum()) == int((dy != 0).sum()) and float(dil.sum()) == pytest.approx(float(dy.sum()), rel=1e-4, abs=1e-2)
    got = _cuda(prev)
    pad = (Cdx + 15) // 16 * 16
    packed = ops_tc.pack_weights_f16(ops_bwd.rot_weights(k, ci_pad=pad))
    ops_bwd.conv3x3_tc_f16_dgrad(dil, packed, got, pad, accumulate=True)
    _close(got, ref, rel=2e-5, name="stride-2 tc dgrad")


@pytest.mark.parametrize("shape", [(2, 7, 16, 192), (1, 14, 32, 128), (1, 28, 64, 32), (2, 5, 9, 16), (1, 1, 2, 32), (1, 13, 21, 36)])
def test_cost_volume_bwd_matches_autograd(P, shape):
    from pwcnet_b200 import ops_bwd
    f0 = torch.from_numpy(_rand(shape, 1)).requires_grad_(True)
    f1 = torch.from_numpy(_rand(shape, 2)).requires_grad_(True)
    cv = O.cost_volume_closed_form(f0, f1, 4)
    g = torch.from_numpy(_rand(tuple(cv.shape), 3))
    cv.backward(g)
    slot = _rand(shape, 6)
    prev0, prev1 = _rand(shape, 7), _rand(shape, 8)
    df0, df1 = _cuda(prev0), _cuda(prev1)
    ops_bwd.cost_volume_bwd(_cuda(g.numpy()), _cuda(cv.detach().numpy()), _cuda(f0.detach().numpy()), _cuda(f1.detach().numpy()),
                            df0, df1, g_f0slot=_cuda(slot), accumulate_f1=True)
    _close(df0, f0.grad + torch.from_numpy(prev0 + slot), name="df0")
    _close(df1, f1.grad + torch.from_numpy(prev1), name="df1 (accumulate)")
    df1b = torch.full(shape, 3.0, device="cuda")
    ops_bwd.cost_volume_bwd(_cuda(g.numpy()), _cuda(cv.detach().numpy()), _cuda(f0.detach().numpy()), _cuda(f1.detach().numpy()),
                            _cuda(prev0), df1b)
    _close(df1b, f1.grad, name="df1 (overwrite)")


@pytest.mark.parametrize("shape,scale,amp", [((2, 14, 32, 128), 1.25, 3.0), ((1, 28, 64, 96), 2.5, 2.0), ((1, 9, 13, 32), 5.0, 4.0),
                                             ((2, 6, 7, 20), 1.0, 30.0)])
def test_warp_bwd_matches_autograd(P, shape, scale, amp):
    """Flows large enough (amp*scale px) that taps clamp at the border; weights stay un-clamped."""
    from pwcnet_b200 import ops_bwd
    B, H, W, C = shape
    x = torch.from_numpy(_rand(shape, 1)).requires_grad_(True)
    flow = torch.from_numpy(_rand((B, H, W, 2), 2, amp)).requires_grad_(True)
    out = O.bilinear_warp(x, flow * scale)
    g = torch.from_numpy(_rand(shape, 3))
    out.backward(g)
    prev = _rand((B, H, W, 2), 5)
    dx = torch.zeros(shape, device="cuda")
    dflow = _cuda(prev)
    ops_bwd.warp_bwd(_cuda(x.detach().numpy()), _cuda(flow.detach().numpy()), _cuda(g.numpy()), dx, dflow, flow_scale=scale)
    _close(dx, x.grad, name="dx")
    _close(dflow, flow.grad + torch.from_numpy(prev), rel=1e-4, name="dflow")
    # nearest: single tap, no flow gradient
    xn = torch.from_numpy(_rand(shape, 1)).requires_grad_(True)
    O.nearest_warp(xn, flow.detach() * scale).backward(g)
    dxn = torch.zeros(shape, device="cuda")
    ops_bwd.warp_bwd(_cuda(xn.detach().numpy()), _cuda(flow.detach().numpy()), _cuda(g.numpy()), dxn, None, flow_scale=scale,
                     warp_type="nearest")
    _close(dxn, xn.grad, name="dx nearest")


@pytest.mark.parametrize("shape,out_hw,mul", [((2, 7, 16, 2), (14, 32), 1.0), ((1, 14, 32, 32), (28, 64), 1.0),
                                              ((1, 5, 6, 3), (20, 24), 20.0), ((1, 4, 5, 2), (7, 11), 1.0)])
def test_resize_bilinear_bwd_matches_autograd(P, shape, out_hw, mul):
    from pwcnet_b200 import ops_bwd
    x = torch.from_numpy(_rand(shape, 1)).requires_grad_(True)
    y = O.resize_bilinear_legacy(x, *out_hw) * mul
    g = torch.from_numpy(_rand(tuple(y.shape), 2))
    y.backward(g)
    prev = _rand(shape, 3)
    dx = _cuda(prev)
    ops_bwd.resize_bilinear_bwd(_cuda(g.numpy()), dx, mul=mul)
    _close(dx, x.grad + torch.from_numpy(prev), name="dx")


@pytest.mark.parametrize("ord", [1, 2])
def test_lploss_bwd_matches_autograd(P, ord):
    from pwcnet_b200 import ops_bwd
    B, H, W, h, w = 2, 64, 128, 8, 16
    gt = torch.from_numpy(_rand((B, H, W, 2), 1, 5.0))
    fs = torch.from_numpy(_rand((B, h, w, 2), 2, 0.3)).requires_grad_(True)
    gd = O.resize_nearest_legacy(gt / 20.0, h, w)
    loss = 0.32 * (O.L2loss(gd, fs) if ord == 2 else O.L1loss(gd, fs))
    loss.backward()
    gfs = torch.empty((B, h, w, 2), device="cuda")
    ops_bwd.lploss_level_bwd(gt.cuda(), _cuda(fs.detach().numpy()), 0.32, gfs, gt_div=20.0, ord=ord)
    _close(gfs, fs.grad, name="gfs")


def test_leaky_add_sumsq_adam(P):
    from pwcnet_b200 import ops_bwd
    y, g0 = _rand((2, 5, 7, 12), 1), _rand((2, 5, 7, 12), 2)
    g = _cuda(g0)
    ops_bwd.leaky_bwd(g, _cuda(y), 0.1)
    assert np.array_equal(g.cpu().numpy(), np.where(y > 0, g0, g0 * np.float32(0.1)).astype(np.float32))
    dst = _cuda(y); ops_bwd.add_(dst, _cuda(g0), 0.5)
    _close(dst, y + 0.5 * g0, name="add_")
    n = 100003
    var, grad = _rand((n,), 3), _rand((n,), 4, 0.01)
    acc = torch.zeros(1, device="cuda")
    ops_bwd.sumsq(_cuda(var), acc, 0.5)
    assert acc.item() == pytest.approx(0.5 * float((var.astype(np.float64) ** 2).sum()), rel=1e-5)
    # three TF-Adam steps with the l2 regulariser folded in (train.py:74,89)
    gamma = 4e-4
    v_gpu, m_gpu, s_gpu = _cuda(var), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    vr, mr, sr = var.copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    lr_dev = torch.zeros(1, device="cuda")
    for t in (1, 2, 3):
        gr = grad * t
        lr_dev.fill_(1e-4 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t))
        ops_bwd.adam_step(v_gpu, _cuda(gr), m_gpu, s_gpu, lr_dev, gamma=gamma)
        vr, mr, sr = O.adam_step_tf(vr, (gr + np.float32(gamma) * vr).astype(np.float32), mr, sr, t, 1e-4)
    np.testing.assert_allclose(v_gpu.cpu().numpy(), vr, rtol=3e-7, atol=2e-7)   # 1 ulp (fma contraction)
    np.testing.assert_allclose(m_gpu.cpu().numpy(), mr, rtol=1e-5, atol=1e-9)


# ----------------------------------------------------------------------------------------- whole network
def _oracle_grads(W, im0, im1, gt, gamma=0.0):
    Wt = {k: torch.from_numpy(v.copy()).requires_grad_(True) for k, v in W.items()}
    total, epe, ff, pyr = O.training_loss(Wt, im0, im1, gt, gamma=gamma)
    total.backward()
    return float(total), float(epe), {k: v.grad.numpy() for k, v in Wt.items()}


@pytest.mark.parametrize("precision", ["fp32", "3xf16"])   # 3xf16: tcgen05 forward AND stride-1 dgrad
def test_network_gradients_match_oracle_autograd(P, precision):
    """All 110 gradient tensors of the multiscale loss at 64x128 with 'hot' weights (flows of several pixels, so
    the warp's flow gradient, border clamping and the three flow routes between levels are all live)."""
    from pwcnet_b200.train import Trainer
    W = O.glorot_weights(7, gain=1.4, bias_scale=0.02)
    im0, im1 = O.synthetic_pair(2, 64, 128, 3, shift=(5, -3))
    gt = np.random.default_rng(1).normal(0, 5, (2, 64, 128, 2)).astype(np.float32)
    ref_total, ref_epe, ref = _oracle_grads(W, im0, im1, gt)
    model = P.PWCDCNet(weights=W, precision=precision)
    tr = Trainer(model)
    tr.forward_backward(im0, im1, gt)
    torch.cuda.synchronize()
    s = tr._scalars.cpu().numpy()
    assert s[0] == pytest.approx(ref_total, rel=2e-5)
    assert s[2] == pytest.approx(ref_epe, abs=1e-3)
    rel = 2e-4 if precision == "fp32" else 1e-3
    worst = 0.0
    for name in model.var_names:
        got, r = tr.grads[name].cpu().numpy(), ref[name]
        scale = float(np.abs(r).max())
        assert scale > 0, f"{name}: reference gradient is identically zero"
        err = float(np.abs(got - r).max()) / scale
        worst = max(worst, err)
        assert err < rel, f"{name}: relative max-abs gradient error {err:.3e}"
    print(f"worst relative gradient error ({precision}): {worst:.2e}")


def test_three_training_steps_match_oracle(P):
    """loss, EPE and the weights after 3 x (forward, backward, TF-Adam with l2 regulariser) vs the oracle loop
    (train.py:65-92: minimise loss + gamma * sum l2_loss(var), lr 1e-4, global_step increments)."""
    from pwcnet_b200.train import Trainer
    W = O.glorot_weights(11, gain=1.2, bias_scale=0.02)
    im0, im1 = O.synthetic_pair(2, 64, 64, 4, shift=(2, 1))
    gt = np.random.default_rng(2).normal(0, 3, (2, 64, 64, 2)).astype(np.float32)
    model = P.PWCDCNet(weights=W, precision="fp32")
    tr = Trainer(model, lr=1e-4, gamma=4e-4)
    Wr = {k: v.copy() for k, v in W.items()}
    M = {k: np.zeros_like(v) for k, v in W.items()}
    V = {k: np.zeros_like(v) for k, v in W.items()}
    for t in (1, 2, 3):
        loss, loss_ms, epe = tr.step(im0, im1, gt)
        ref_total, ref_epe, g = _oracle_grads(Wr, im0, im1, gt, gamma=4e-4)
        assert loss.item() == pytest.approx(ref_total, rel=5e-5), f"step {t}"
        assert epe.item() == pytest.approx(ref_epe, abs=1e-3)
        for k in Wr:
            Wr[k], M[k], V[k] = O.adam_step_tf(Wr[k], g[k], M[k], V[k], t, O.piecewise_lr(t - 1, 1e-4))
    assert tr.global_step == 3
    sd = model.state_dict()
    for k in Wr:
        # Adam's first steps move every weight by ~lr regardless of gradient size; sign flips of tiny gradients
        # are the only way to differ, bounded by 2*lr per step
        np.testing.assert_allclose(sd[k], Wr[k], rtol=0, atol=2e-5, err_msg=k)
    moved = max(float(np.abs(sd[k] - W[k]).max()) for k in W)
    assert moved > 1e-4
    # the derived kernels were refreshed: a fresh model with the updated weights gives the same flow
    ff, _ = model(im0, im1)
    ff2, _ = P.PWCDCNet(weights=sd, precision="fp32")(im0, im1)
    assert torch.equal(ff, ff2)


def test_train_stream_matches_synchronous_steps(P):
    """TrainStream (staged H2D on a copy stream, batch i+1 in flight during step i) trains exactly like
    Trainer.step on the same host batches, including when a staging set is reused."""
    from pwcnet_b200.train import Trainer
    W = O.glorot_weights(5, gain=1.2, bias_scale=0.02)
    rng = np.random.default_rng(9)
    batches = []
    for i in range(4):
        im0, im1 = O.synthetic_pair(2, 64, 64, 20 + i, shift=(1, -1))
        batches.append((torch.from_numpy(im0).pin_memory(), torch.from_numpy(im1).pin_memory(),
                        torch.from_numpy(rng.normal(0, 3, (2, 64, 64, 2)).astype(np.float32)).pin_memory()))
    ref_tr = Trainer(P.PWCDCNet(weights=W))
    ref = [[float(v) for v in ref_tr.step(*b)] for b in batches]
    tr = Trainer(P.PWCDCNet(weights=W))
    ts = P.TrainStream(tr, depth=2)
    with pytest.raises(RuntimeError):
        ts.step()
    got = []
    ts.submit(*batches[0])
    for i in range(4):
        if i + 1 < 4:
            ts.submit(*batches[i + 1])
        got.append([float(v) for v in ts.step()])
    assert ts.pending() == 0 and tr.global_step == 4
    np.testing.assert_allclose(got, ref, rtol=1e-4)          # fp32 atomics in the wgrad kernels: not bit-reproducible
    assert float((tr.model.flat - ref_tr.model.flat).abs().max()) < 1e-4
    ts.submit(*batches[0]); ts.submit(*batches[1])
    with pytest.raises(RuntimeError):
        ts.submit(*batches[2])


def test_batched_weight_packs_equal_per_layer_packs(P):
    """The one-launch packs (dgrad kernels before a backward pass, forward kernels after the Adam update) write exactly
    the bytes of the per-layer rot_weights + pack_weights_f16 calls, and the gradient does not depend on which ran."""
    from pwcnet_b200 import ops_bwd, ops_tc
    from pwcnet_b200.train import Trainer
    W = O.glorot_weights(6, gain=1.2, bias_scale=0.02)
    im0, im1 = O.synthetic_pair(2, 64, 128, 31, shift=(2, -1))
    gt = np.random.default_rng(3).normal(0, 3, (2, 64, 128, 2)).astype(np.float32)
    tr = Trainer(P.PWCDCNet(weights=W, cv_pipeline="default"))
    m = tr.model
    tr.forward_backward(im0, im1, gt)          # first pass: layer-by-layer packs, records the parts
    g1 = tr.grad_flat.clone()
    tr.forward_backward(im0, im1, gt)          # second pass: one batched launch
    g2 = tr.grad_flat.clone()
    assert len(tr._dgrad_jobs) > 30 and tr._dgrad_batched == set(tr._parts)
    assert float((g1 - g2).abs().max()) < 1e-5 * float(g1.abs().max())
    for scope, parts in tr._parts.items():
        for c0, cnt, pad, rot, packed in parts:
            ref = ops_tc.pack_weights_f16(ops_bwd.rot_weights(m._k[scope], ci_begin=c0, ci_count=cnt, ci_pad=pad))
            assert torch.equal(ref, packed), scope
    tr.step(im0, im1, gt)                      # Adam update -> refresh_derived -> batched forward packs
    assert len(m._pack_keys) > 40
    for key in m._pack_keys:
        scope = key.split("#")[0]
        if key.endswith("#s2d"):               # stride-2 conv as a 2x2 conv: the pack's source is the re-indexed kernel
            assert torch.equal(m._k_s2d[scope], ops_tc.s2d_reindex(m._k[scope])), key
            src = m._k_s2d[scope]
        else:
            src = m._head_k[scope] if key.endswith("#head") else m._k[scope]
        assert torch.equal(ops_tc.pack_weights_f16(src), m._packed[key]), key
    ff, _ = m(im0, im1)
    ff2, _ = P.PWCDCNet(weights=m.state_dict(), cv_pipeline="default")(im0, im1)      # the trainer's model runs the default pipeline
    assert torch.equal(ff, ff2)
    ff3, _ = P.PWCDCNet(weights=m.state_dict())(im0, im1)                              # split cost-volume pipeline: fp32-class, not bit-equal
    assert float((ff3 - ff).abs().max()) < 1e-4


def test_trainer_rejects_unsupported_models(P):
    from pwcnet_b200.train import Trainer
    with pytest.raises(P.PwcError):
        Trainer(P.PWCDCNet(fuse_warp=True))
    with pytest.raises(P.PwcError):
        Trainer(P.PWCDCNet(precision="cudnn"))


@pytest.mark.parametrize("precision", ["3xf16", "fp32"])     # tcgen05 dgrad epilogues / CUDA-core dgrad kernels
@pytest.mark.parametrize("use_dc", [False, True])
@pytest.mark.parametrize("shape", [(2, 64, 128), (1, 192, 320)])
def test_uncleared_gradient_buffers_are_fully_overwritten(P, use_dc, shape, precision):
    """`Trainer.backward` clears only the activation-gradient buffers whose first writer accumulates (`_grad_buffers`);
    the others must be overwritten completely by their one dgrad.  Poison them with NaN between two backward passes over
    the same batch: every gradient stays finite and unchanged (flat-slot and row-tile dgrad epilogues, plain and dense stacks)."""
    from pwcnet_b200.train import Trainer
    B, H, W_ = shape
    W = O.glorot_weights(7, gain=1.2, bias_scale=0.02, use_dc=use_dc)
    im0, im1 = O.synthetic_pair(B, H, W_, 3, shift=(3, -2))
    gt = np.random.default_rng(2).normal(0, 4, (B, H, W_, 2)).astype(np.float32)
    model = P.PWCDCNet(weights=W, use_dc=use_dc, precision=precision)
    tr = Trainer(model)
    tr.forward_backward(im0, im1, gt)
    torch.cuda.synchronize()
    first = tr.grad_flat.clone()
    (g,) = tr._gbufs.values()
    assert 0 < g.n_clear < g.flat.numel()
    g.flat[g.n_clear:].fill_(float("nan"))
    tr.forward_backward(im0, im1, gt)
    torch.cuda.synchronize()
    assert torch.isfinite(tr.grad_flat).all()
    # (reduce-add epilogues and split wgrad sums are not order-deterministic: equal up to fp32 summation order)
    assert float((tr.grad_flat - first).abs().max()) <= 1e-5 * float(first.abs().max())


@pytest.mark.parametrize("precision", ["fp32", "3xf16"])
def test_use_dc_network_gradients_match_oracle_autograd(P, precision):
    """`--use-dc` (train.py:203-207, modules.py:269-270): all 110 gradient tensors of the densely connected estimator
    network vs torch autograd over the oracle, and one optimisation step."""
    from pwcnet_b200.train import Trainer
    W = O.glorot_weights(5, gain=1.0, bias_scale=0.02, use_dc=True)
    im0, im1 = O.synthetic_pair(2, 64, 128, 4, shift=(-4, 2))
    gt = np.random.default_rng(1).normal(0, 5, (2, 64, 128, 2)).astype(np.float32)
    Wt = {k: torch.from_numpy(v.copy()).requires_grad_(True) for k, v in W.items()}
    total, epe, ff, pyr = O.training_loss(Wt, im0, im1, gt, gamma=0.0, use_dc=True)
    total.backward()
    model = P.PWCDCNet(weights=W, use_dc=True, precision=precision)
    tr = Trainer(model)
    tr.forward_backward(im0, im1, gt)
    torch.cuda.synchronize()
    s = tr._scalars.cpu().numpy()
    assert s[0] == pytest.approx(float(total), rel=2e-5)
    # fp32 proves the orchestration (2e-4).  With the tensor-core forward the dense stack feeds every pre-activation to
    # up to seven consumers: a handful of leaky masks of near-zero pre-activations flip relative to the fp32 forward and
    # move single gradient tensors by a few 1e-3 of their maximum (identical with CUDA-core dgrad / wgrad: tools/dc_grad_dbg.py)
    rel = 2e-4 if precision == "fp32" else 1e-2
    worst = 0.0
    for name in model.var_names:
        got, r = tr.grads[name].cpu().numpy(), Wt[name].grad.numpy()
        scale = float(np.abs(r).max())
        assert scale > 0, name
        err = float(np.abs(got - r).max()) / scale
        worst = max(worst, err)
        assert err < rel, f"{name}: relative max-abs gradient error {err:.3e}"
    print(f"use_dc worst relative gradient error ({precision}): {worst:.2e}")
    loss, _, _ = tr.step(im0, im1, gt)
    assert np.isfinite(loss.item()) and tr.global_step == 1


def test_trainer_state_dict_roundtrip_resumes_identically(P, tmp_path):
    """Saver semantics (train.py:95-99,164-166): weights + Adam slots + global_step under the reference's names;
    a trainer restored from the file continues bit-for-bit like the original."""
    from pwcnet_b200.train import Trainer
    W = O.glorot_weights(5, gain=1.2, bias_scale=0.02)
    im0, im1 = O.synthetic_pair(1, 64, 64, 9, shift=(1, 2))
    gt = np.random.default_rng(3).normal(0, 3, (1, 64, 64, 2)).astype(np.float32)
    tr = Trainer(P.PWCDCNet(weights=W, precision="fp32"))
    tr.step(im0, im1, gt); tr.step(im0, im1, gt)
    path = str(tmp_path / "model_2.ckpt")
    tr.save(path)                                   # a TF checkpoint bundle, as train.py:166 writes
    from pwcnet_b200.checkpoint import load_all
    sd = load_all(path)
    assert int(sd["Variable"]) == 2 and "pwcdcnet/context/conv2d_6/kernel/Adam_1" in sd and len(sd) == 110 * 3 + 3
    assert sd["Variable"].dtype == np.int32 and sd["Variable"].shape == ()
    tr2 = Trainer(P.PWCDCNet(precision="fp32"))
    tr2.load_state_dict(path)                       # resume by path (train.py:97-99)
    # ... and inference by path (test.py:40-42)
    ffa, _ = P.PWCDCNet(weights=path, precision="fp32")(im0, im1)
    ffb, _ = tr.model(im0, im1)
    assert torch.equal(ffa, ffb)
    assert tr2.global_step == 2
    tr.step(im0, im1, gt); tr2.step(im0, im1, gt)
    # wgrad accumulates with float atomics: summation order is not fixed, so compare to 1e-6 instead of bit-for-bit
    np.testing.assert_allclose(tr2.model.flat.cpu().numpy(), tr.model.flat.cpu().numpy(), rtol=0, atol=2e-6)
