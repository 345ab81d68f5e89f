"""GPU-vs-oracle parity AT THE BENCHMARKED SHAPES (VERDICT r1 "What's weak" 1, 2, 4): whole-network forward at
448x1024 (BASELINE configs 2-4), one training step at 384x1024 (config 5), and the reference's trained checkpoint
(tests/golden/model_250_weights.npz + trained_model250_*.npz, written by oracle/make_golden.py from the reference's own
GraphDef + checkpoint).  At these sizes the halo conv runs in its 128-pixel row-tile mode, the pyramid in its full-width
first layers and the context net with dilation 16 on 112x256 -- the code paths bench.py times.

Tolerances (north_star): final flow <= 1e-3 max-abs; every pyramid flow <= 5e-5 (= 1e-3 / 20, GT/20 units).
Measured errors are printed (pytest -s) and quoted in DESIGN.md section 4."""
import os

import numpy as np
import pytest
import torch

from oracle import pwc_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def P():
    import pwcnet_b200 as P
    assert torch.cuda.is_available()
    return P


@pytest.fixture(scope="module")
def trained():
    return dict(np.load(os.path.join(GOLD, "model_250_weights.npz")))


def _report(tag, ff, pyr, rff, rpyr):
    e_ff = float(np.abs(ff.cpu().numpy() - np.asarray(rff)).max())
    e_py = [float(np.abs(a.cpu().numpy() - np.asarray(b)).max()) for a, b in zip(pyr, rpyr)]
    print(f"[parity {tag}] max|flow| {float(np.abs(np.asarray(rff)).max()):.2f} px  final-flow max-abs err {e_ff:.3e}  "
          f"pyramid errs {' '.join(f'{e:.1e}' for e in e_py)}")
    return e_ff, e_py


CASES_448 = {
    "glorot2_random": dict(weights=("glorot", 2, 1.0, 0.0), pair=("random", 0, None)),
    "hot7_shift": dict(weights=("glorot", 7, 1.4, 0.02), pair=("random", 3, (5, -3))),
    "hot7_texture12": dict(weights=("glorot", 7, 1.4, 0.02), pair=("texture", 5, 12.0)),
}


def _inputs(case, H, W):
    kind, seed, gain, bias = case["weights"]
    Wt = O.glorot_weights(seed, gain=gain, bias_scale=bias)
    pk, pseed, parg = case["pair"]
    if pk == "random":
        im0, im1 = O.synthetic_pair(1, H, W, pseed, shift=parg)
    else:
        im0, im1, _ = O.synthetic_textured_pair(1, H, W, pseed, parg)
    return Wt, im0, im1


@pytest.mark.parametrize("name", list(CASES_448))
def test_forward_448x1024_vs_oracle(P, name):
    """PWCDCNet (default 3xf16 tcgen05 path, CUDA graph) vs oracle.pwcdcnet_forward at BASELINE's 448x1024, B=1."""
    Wt, im0, im1 = _inputs(CASES_448[name], 448, 1024)
    model = P.PWCDCNet(weights=Wt)
    ff, pyr = model(im0, im1)
    rff, rpyr = O.pwcdcnet_forward(Wt, im0, im1)
    e_ff, e_py = _report(f"448x1024 {name}", ff, pyr, rff.numpy(), [p.numpy() for p in rpyr])
    assert e_ff <= 1e-3
    assert max(e_py) <= 5e-5
    # the captured graph replays to the same bits, and a batch of 2 reproduces the single pair
    ff = ff.clone()
    ff2, _ = model(im0, im1)
    assert torch.equal(ff, ff2)
    fb, _ = model(np.concatenate([im0, im0]), np.concatenate([im1, im1]))
    assert torch.equal(fb[0], ff[0]) and torch.equal(fb[1], ff[0])


@pytest.mark.parametrize("hw", [(64, 128), (448, 1024)])
def test_trained_checkpoint_vs_reference_graph(P, trained, hw):
    """The reference's trained model_250 weights on the GPU path vs the reference's GraphDef outputs (golden fixture)."""
    H, W = hw
    gold = np.load(os.path.join(GOLD, f"trained_model250_{H}x{W}.npz"))
    im0, im1, flow = O.synthetic_textured_pair(1, H, W, int(gold["seed"]), float(gold["max_disp"]))
    model = P.PWCDCNet(weights=trained)
    ff, pyr = model(im0, im1)
    e_ff, e_py = _report(f"trained {H}x{W}", ff, pyr, gold["flows_final"], [gold[f"pyr{l}"] for l in range(5)])
    assert e_ff <= 1e-3
    assert max(e_py) <= 5e-5
    epe = P.EPE(torch.from_numpy(flow).cuda(), ff).item()
    assert abs(epe - float(gold["epe"])) < 1e-3
    loss = P.multiscale_loss(torch.from_numpy(flow).cuda(), pyr, O.DEFAULT_LOSS_WEIGHTS).item()
    assert loss == pytest.approx(float(gold["loss"]), rel=1e-4)
    if H == 448:
        # the trained net actually recovers the synthetic motion (sub-pixel EPE): a semantic end-to-end check
        assert epe < 1.0


def test_trained_checkpoint_all_precisions_full_size(P, trained):
    """fp32 CUDA-core, 3xf16 and 3xtf32 agree with the reference graph at 448x1024 with trained weights."""
    gold = np.load(os.path.join(GOLD, "trained_model250_448x1024.npz"))
    im0, im1, _ = O.synthetic_textured_pair(1, 448, 1024, int(gold["seed"]), float(gold["max_disp"]))
    for precision in ("fp32", "3xtf32"):
        ff, pyr = P.PWCDCNet(weights=trained, precision=precision)(im0, im1)
        e_ff, e_py = _report(f"trained 448x1024 {precision}", ff, pyr, gold["flows_final"], [gold[f"pyr{l}"] for l in range(5)])
        assert e_ff <= 1e-3 and max(e_py) <= 5e-5


def _oracle_grads(W, im0, im1, gt, gamma=0.0):
    Wt = {k: torch.from_numpy(v.copy()).requires_grad_(True) for k, v in W.items()}
    total, epe, ff, pyr = O.training_loss(Wt, im0, im1, gt, gamma=gamma)
    total.backward()
    return float(total), float(epe), {k: v.grad.numpy() for k, v in Wt.items()}


def test_training_step_384x1024_vs_oracle_autograd(P):
    """BASELINE config 5's shape (Sintel 436x1024 cropped to /64 = 384x1024), B=1: loss, EPE and all 110 gradient
    tensors of the default (3xf16 tcgen05 forward + dgrad + wgrad) training path vs torch autograd over the oracle."""
    from pwcnet_b200.train import Trainer
    W = O.glorot_weights(7, gain=1.4, bias_scale=0.02)
    im0, im1, flow = O.synthetic_textured_pair(1, 384, 1024, 9, 10.0)
    gt = (flow + np.random.default_rng(1).normal(0, 2, flow.shape)).astype(np.float32)
    ref_total, ref_epe, ref = _oracle_grads(W, im0, im1, gt)
    model = P.PWCDCNet(weights=W)
    tr = Trainer(model)
    tr.forward_backward(im0, im1, gt)
    torch.cuda.synchronize()
    s = tr._scalars.cpu().numpy()
    assert s[0] == pytest.approx(ref_total, rel=5e-5)
    assert s[2] == pytest.approx(ref_epe, abs=1e-3)
    worst, worst_name = 0.0, ""
    for name in model.var_names:
        got, r = tr.grads[name].cpu().numpy(), ref[name]
        scale = float(np.abs(r).max())
        assert scale > 0, f"{name}: reference gradient is identically zero"
        err = float(np.abs(got - r).max()) / scale
        if err > worst:
            worst, worst_name = err, name
        assert err < 1e-3, f"{name}: relative max-abs gradient error {err:.3e}"
    print(f"[parity train 384x1024] loss {s[0]:.6f} vs {ref_total:.6f}  EPE {s[2]:.5f} vs {ref_epe:.5f}  "
          f"worst relative gradient error {worst:.2e} ({worst_name})")


def test_fp16_range_guard_fails_loudly(P):
    """VERDICT r1 weak point 3: the 3 x fp16 split needs |x| < 65504.  Activations of ~1e5 (first conv scaled up) must not
    silently turn into garbage: the forward counts non-finite flow values and check_finite / the next call / collect raise;
    the fp32-operand precisions handle the same weights."""
    W = O.glorot_weights(2)
    W = {k: v.copy() for k, v in W.items()}
    W["pwcdcnet/fp_extractor/conv2d/kernel"] *= 4e5          # level-1 activations of ~1e5, beyond fp16
    W["pwcdcnet/fp_extractor/conv2d_1/kernel"] *= 1e-5       # ... scaled back by the next layer: fp32 arithmetic is unaffected
    im0, im1 = O.synthetic_pair(1, 64, 192, 0)      # level-1 rows of 96 pixels: the 16 -> 16 layer runs on the fp16 halo kernel
    rff, _ = O.pwcdcnet_forward(W, im0, im1)
    assert float(O.pyramid_extractor(torch.from_numpy(im0), W)[-1].abs().max()) < 1e3      # the oracle stays finite and small
    model = P.PWCDCNet(weights=W)                               # default: 3xf16
    ff, _ = model(im0, im1)
    with pytest.raises(P.PwcError, match="non-finite"):
        model.check_finite()
    model(im0, im1)
    torch.cuda.synchronize()
    with pytest.raises(P.PwcError, match="non-finite"):
        model(im0, im1)                                         # the previous call's overflow surfaces on the next call
    model.check_finite()                                        # consumed
    st = P.InferenceStream(model, depth=2)
    with pytest.raises(P.PwcError, match="non-finite"):
        st.collect(st.submit(im0, im1))
    for precision in ("fp32", "3xtf32"):
        m2 = P.PWCDCNet(weights=W, precision=precision)
        ff2, _ = m2(im0, im1)
        m2.check_finite()
        np.testing.assert_allclose(ff2.cpu().numpy(), rff.numpy(), atol=1e-3, rtol=0)
    ok = P.PWCDCNet(weights=O.glorot_weights(2))
    ok(im0, im1)
    ok.check_finite()
