"""The inference CLIs (reference test.py / test_continuous.py equivalents) end to end on the GPU: generated PNG frames in,
Middlebury `.flo` files out, compared with the model called directly on the reference's float feed."""
import os

import numpy as np
import pytest
import torch

from oracle import pwc_oracle as O

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


def _frames(tmp_path, n, H, W, seed=0):
    im0, _, _ = O.synthetic_textured_pair(1, H, W, seed, 3.0)
    base = (im0[0] * 255).astype(np.uint8)
    paths = []
    d = tmp_path / "seq" / "clip"
    d.mkdir(parents=True)
    for i in range(n):
        frame = np.roll(base, (i, 2 * i), axis=(0, 1))
        p = str(d / f"frame_{i:04d}.png")
        cv2.imwrite(p, cv2.cvtColor(frame, cv2.COLOR_RGB2BGR))
        paths.append(p)
    return paths, [np.roll(base, (i, 2 * i), axis=(0, 1)) for i in range(n)]


def test_infer_cli_two_images_to_flo(tmp_path):
    import pwcnet_b200 as P
    from pwcnet_b200 import infer, flow_io
    paths, frames = _frames(tmp_path, 2, 70, 140)            # 70x140 -> factor_crop -> 64x128
    out = str(tmp_path / "flow.flo")
    infer.main(["--input_images", paths[0], paths[1], "--out", out])
    flo = flow_io.load_flow(out)
    assert flo.shape == (64, 128, 2)
    feed = [(flow_io.factor_crop(f)[None] / 255.0).astype(np.float32) for f in frames]
    ref, _ = P.PWCDCNet()(feed[0], feed[1])                   # same default (glorot seed 0) weights as the CLI's model
    np.testing.assert_array_equal(flo, ref[0].cpu().numpy())


def test_infer_continuous_cli_sliding_pairs(tmp_path):
    import pwcnet_b200 as P
    from pwcnet_b200 import infer_continuous, flow_io
    paths, frames = _frames(tmp_path, 4, 64, 128, seed=1)
    out_dir = str(tmp_path / "figs")
    written = infer_continuous.main(["-i", *paths, "--out_dir", out_dir])
    assert [os.path.basename(w) for w in written] == ["frame_0000.flo", "frame_0001.flo", "frame_0002.flo"]
    assert all(os.path.dirname(w).endswith(os.path.join("figs", "clip")) for w in written)      # <out_dir>/<dname>/<fname>.flo
    model = P.PWCDCNet()
    for i, w in enumerate(written):
        ref, _ = model((frames[i][None] / 255.0).astype(np.float32), (frames[i + 1][None] / 255.0).astype(np.float32))
        np.testing.assert_array_equal(flow_io.load_flow(w), ref[0].cpu().numpy())
    with pytest.raises(ValueError):
        infer_continuous.main(["-i", paths[0]])


def test_train_cli_runs_epochs_summaries_checkpoints(tmp_path, monkeypatch):
    """The reference's train.py loop (train.py:118-172) end to end on a tiny synthetic Sintel tree: DataLoader -> uint8
    batches -> TrainStream -> summaries, validation, per-epoch TF-bundle checkpoints, ExperimentSaver; then --resume."""
    import json
    import random
    from pwcnet_b200 import train_cli, checkpoint
    from pwcnet_b200.flow_io import save_flow
    rng = np.random.default_rng(0)
    root = tmp_path / "sintel"
    for s in ("alley_1", "cave_2"):
        for i in range(1, 7):
            (root / "training/clean" / s).mkdir(parents=True, exist_ok=True)
            (root / "training/flow" / s).mkdir(parents=True, exist_ok=True)
            cv2.imwrite(str(root / "training/clean" / s / f"frame_{i:04d}.png"), rng.integers(0, 256, (72, 136, 3), dtype=np.uint8))
            if i < 6:
                save_flow(str(root / "training/flow" / s / f"frame_{i:04d}.flo"), rng.normal(0, 2, (72, 136, 2)).astype(np.float32))
    monkeypatch.chdir(tmp_path)
    random.seed(0)
    argv = ["-dd", str(root), "-e", "2", "-b", "2", "-nw", "0", "--crop_shape", "64", "128", "--summary_every", "2"]
    logdir = train_cli.main(argv)
    # 10 pairs -> 9 train / 1 val; 4 batches of 2 per epoch (drop_last), 2 epochs = 8 steps
    tr = [json.loads(l) for l in open(os.path.join(logdir, "train", "scalars.jsonl"))]
    assert [r["step"] for r in tr] == [2, 4, 6, 8] and all(np.isfinite(r["loss/pwc"]) and r["EPE/source"] > 0 for r in tr)
    assert not os.path.exists("model") and os.path.exists(os.path.join(logdir, "config.json"))       # moved by ExperimentSaver
    ck = os.path.join(logdir, "model", "model_2.ckpt")
    sd = checkpoint.load_all(ck)
    assert int(sd["Variable"]) == 8 and len(sd) == 333
    # resume: global_step continues from the checkpoint (train.py:97-99)
    os.makedirs("keep", exist_ok=True)
    logdir2 = train_cli.main(argv[:3] + ["1"] + argv[4:] + ["-r", ck, "--max_steps", "9"])
    sd2 = checkpoint.load_all(os.path.join(logdir2, "model", "model_1.ckpt"))
    assert int(sd2["Variable"]) == 9
    assert float(np.abs(sd2["pwcdcnet/context/conv2d_6/kernel"] - sd["pwcdcnet/context/conv2d_6/kernel"]).max()) > 0
    with pytest.raises(NotImplementedError):
        train_cli.main(argv + ["--loss", "robust"])
