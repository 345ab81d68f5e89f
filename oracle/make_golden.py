"""Generates tests/golden/*.npz from the oracle (run in the build container: `python -m oracle.make_golden`).
The fixtures freeze the oracle's outputs so that later edits to it cannot drift unnoticed, and give the
GPU tests a target that does not depend on executing the oracle on the GPU box."""
import os

import numpy as np
import torch

from oracle import pwc_oracle as O

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def config1():
    W = O.glorot_weights(2)
    im0, im1 = O.synthetic_pair(1, 64, 128, 0)
    ff, pyr = O.pwcdcnet_forward(W, im0, im1)
    gt = np.random.default_rng(1).normal(0, 5, (1, 64, 128, 2)).astype(np.float32)
    d = {"flows_final": ff.numpy(), "epe": np.float32(O.EPE(torch.from_numpy(gt), ff).item()),
         "loss": np.float32(O.multiscale_loss(torch.from_numpy(gt), pyr, O.DEFAULT_LOSS_WEIGHTS).item())}
    for l, p in enumerate(pyr):
        d[f"pyr{l}"] = p.numpy()
    np.savez_compressed(os.path.join(OUT, "config1_glorot_seed2.npz"), **d)


def hot_weights_case():
    """'Hot' seeded weights (gain 1.4, random biases): flows of several pixels, so warping, border
    clamping and the residual path are exercised without shipping a trained checkpoint."""
    W = O.glorot_weights(7, gain=1.4, bias_scale=0.02)
    im0, im1 = O.synthetic_pair(2, 64, 128, 3, shift=(5, -3))
    ff, pyr = O.pwcdcnet_forward(W, im0, im1)
    d = {"flows_final": ff.numpy()}
    for l, p in enumerate(pyr):
        d[f"pyr{l}"] = p.numpy()
    np.savez_compressed(os.path.join(OUT, "hot_seed7_64x128.npz"), **d)
    return ff


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    config1()
    ff = hot_weights_case()
    print("hot flows_final abs max", ff.abs().max().item(), "mean abs", ff.abs().mean().item())
