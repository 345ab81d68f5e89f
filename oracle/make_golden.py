"""Generates tests/golden/*.npz (run in the build container: `python -m oracle.make_golden`).

The fixtures are produced by executing the REFERENCE's own saved TensorFlow-1.8 GraphDef
(/root/reference/model_250epochs_ft_Final/model_250.ckpt.meta, the graph train.py built from model.py /
modules.py / losses.py) with `oracle/tf_graph_interp.py` on seeded synthetic weights and inputs -- the
closest thing to running the reference that this TensorFlow-less environment allows.  The script also
checks that oracle/pwc_oracle.py agrees with that graph before writing.  The GPU box has no
/root/reference: tests there only read the committed .npz files."""
import os

import numpy as np
import torch

from oracle import pwc_oracle as O
from oracle import tf_graph_interp as G

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
META = "/root/reference/model_250epochs_ft_Final/model_250.ckpt.meta"


def _case(fname, W, im0, im1, gt, tol):
    ff, pyr, (msl, total, epe) = G.run_reference_graph(META, W, np.stack([im0, im1], 1),
                                                       [G.MULTISCALE_LOSS, G.TOTAL_LOSS, G.EPE], flows_gt=gt)
    loss, oepe, off, opyr = O.training_loss({k: torch.from_numpy(v) for k, v in W.items()}, im0, im1, gt)
    err = float(np.abs(ff - off.numpy()).max())
    assert err < tol, f"oracle disagrees with the reference GraphDef: {err}"
    for a, b in zip(pyr, opyr):
        assert float(np.abs(a - b.numpy()).max()) < tol / 20
    assert abs(float(total) - loss.item()) < 1e-4 * abs(float(total)) and abs(float(epe) - oepe.item()) < 1e-4
    d = {"flows_final": ff, "epe": np.float32(epe), "loss": np.float32(msl), "total_loss": np.float32(total),
         "source": np.array("reference GraphDef model_250.ckpt.meta executed by oracle/tf_graph_interp.py")}
    for l, p in enumerate(pyr):
        d[f"pyr{l}"] = p
    np.savez_compressed(os.path.join(OUT, fname), **d)
    print(f"{fname}: max|flow| {np.abs(ff).max():.3f}  oracle-vs-graph max-abs {err:.2e}  loss {float(msl):.5f}  EPE {float(epe):.5f}")


CKPT = "/root/reference/model_250epochs_ft_Final/model_250.ckpt"


def _trained_cases():
    """Trained-weights fixtures: the reference's GraphDef executed with the reference's
    own model_250 checkpoint on a shifted-texture pair at 64x128 and at BASELINE's 448x1024.  The 110 checkpoint tensors
    are committed as fp32 (tests/golden/model_250_weights.npz, ~18 MB: the GPU box has no /root/reference)."""
    from pwcnet_b200.checkpoint import load_checkpoint
    W = load_checkpoint(CKPT)
    assert len(W) == 110 and sum(v.size for v in W.values()) == 5029868
    np.savez_compressed(os.path.join(OUT, "model_250_weights.npz"), **W)
    for (H, Wd, seed, disp) in ((64, 128, 5, 4.0), (448, 1024, 5, 12.0)):
        im0, im1, flow = O.synthetic_textured_pair(1, H, Wd, seed, disp)
        ff, pyr, (msl, epe) = G.run_reference_graph(CKPT + ".meta", W, np.stack([im0, im1], 1), [G.MULTISCALE_LOSS, G.EPE],
                                                    flows_gt=flow)
        off, opyr = O.pwcdcnet_forward(W, im0, im1)
        err = float(np.abs(ff - off.numpy()).max())
        assert err < 2e-5, f"oracle disagrees with the reference GraphDef: {err}"
        for a, b in zip(pyr, opyr):
            assert float(np.abs(a - b.numpy()).max()) < 2e-6
        d = {"flows_final": ff, "epe": np.float32(epe), "loss": np.float32(msl), "seed": np.int32(seed), "max_disp": np.float32(disp),
             "source": np.array("reference GraphDef model_250.ckpt.meta + model_250.ckpt executed by oracle/tf_graph_interp.py; "
                                "inputs = oracle.synthetic_textured_pair(1,H,W,seed,max_disp), flows_gt = its displacement field")}
        for l, p in enumerate(pyr):
            d[f"pyr{l}"] = p
        fname = f"trained_model250_{H}x{Wd}.npz"
        np.savez_compressed(os.path.join(OUT, fname), **d)
        print(f"{fname}: max|flow| {np.abs(ff).max():.3f}  EPE vs the true displacement {float(epe):.4f} px  "
              f"oracle-vs-graph max-abs {err:.2e}")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    _trained_cases()
    gt1 = np.random.default_rng(1).normal(0, 5, (1, 64, 128, 2)).astype(np.float32)
    # BASELINE config 1: one 64x128 pair, glorot weights (seed 2)
    _case("config1_glorot_seed2.npz", O.glorot_weights(2), *O.synthetic_pair(1, 64, 128, 0), gt1, 1e-5)
    # 'hot' seeded weights (gain 1.4, random biases): flows up to ~14 px, so warping, border clamping and
    # the residual paths are exercised without shipping a trained checkpoint
    gt2 = np.random.default_rng(1).normal(0, 5, (2, 64, 128, 2)).astype(np.float32)
    _case("hot_seed7_64x128.npz", O.glorot_weights(7, gain=1.4, bias_scale=0.02),
          *O.synthetic_pair(2, 64, 128, 3, shift=(5, -3)), gt2, 2e-4)
