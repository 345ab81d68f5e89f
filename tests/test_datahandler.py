"""Data pipeline (pwcnet_b200/datahandler.py; reference datahandler/flow.py + utils.py) on small synthetic dataset trees.
CPU only.  Each behaviour is checked against what the reference's code does on the same files (restated inline)."""
import os
import random

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from pwcnet_b200 import datahandler as D
from pwcnet_b200.flow_io import save_flow


def _png(path, arr):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    assert cv2.imwrite(path, cv2.cvtColor(arr, cv2.COLOR_RGB2BGR))


def _sintel_tree(root, scenes=("alley_1", "cave_2"), frames=4, h=40, w=56, seed=0):
    rng = np.random.default_rng(seed)
    truth = {}
    for mode in ("clean", "final"):
        for s in scenes:
            for i in range(1, frames + 1):
                img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
                p = f"{root}/training/{mode}/{s}/frame_{i:04d}.png"
                _png(p, img)
                truth[p] = img
    for s in scenes:
        for i in range(1, frames):
            fl = rng.normal(0, 3, (h, w, 2)).astype(np.float32)
            p = f"{root}/training/flow/{s}/frame_{i:04d}.flo"
            os.makedirs(os.path.dirname(p), exist_ok=True)
            save_flow(p, fl)
            truth[p] = fl
    return truth


def test_helpers_match_reference_semantics():
    assert list(D.window([1, 2, 3, 4], 2)) == [(1, 2), (2, 3), (3, 4)]
    assert list(D.window([1], 2)) == []
    assert D.get_size(None, (384, 448), None, None) == (384, 448)
    assert D.get_size((436, 1024), (384, 448), (200, 300), None) == (200, 300)           # resize > crop > origin
    assert D.get_size((436, 1024), None, None, (0.5, 0.25)) == (218.0, 256.0)
    with pytest.raises(ValueError):
        D.get_size()
    img = np.arange(10 * 12).reshape(10, 12)
    c = D.StaticCenterCrop((10, 12), (4, 6))
    np.testing.assert_array_equal(c(img), img[3:7, 3:9])
    c = D.StaticCenterCrop((11, 13), (4, 6))                                              # odd sizes: floor on both ends
    np.testing.assert_array_equal(c(img[:, :]), img[(11 - 4) // 2:(11 + 4) // 2, (13 - 6) // 2:(13 + 6) // 2])
    random.seed(7)
    r = D.StaticRandomCrop((10, 12), (4, 6))
    random.seed(7)
    h1, w1 = random.randint(0, 6), random.randint(0, 6)                                   # rows first, then columns (utils.py:10-11)
    assert (r.h1, r.w1) == (h1, w1)
    np.testing.assert_array_equal(r(img), img[h1:h1 + 4, w1:w1 + 6])
    fl = np.stack([np.full((4, 6), 2.0, np.float32), np.full((4, 6), -3.0, np.float32)], -1)
    out = D.resize_flow(fl, (8, 9))
    assert out.shape == (8, 9, 2) and out.dtype == np.float32
    np.testing.assert_allclose(out[..., 0], 2.0 * 9 / 6, rtol=1e-6)                       # u scales with the width ratio
    np.testing.assert_allclose(out[..., 1], -3.0 * 8 / 4, rtol=1e-6)
    out = D.rescale_flow(fl, (0.5, 0.5))
    assert out.shape == (2, 3, 2)
    np.testing.assert_allclose(out[..., 0], 1.0, rtol=1e-6)
    with pytest.raises(ValueError):
        D.resize_flow(fl[..., 0], (8, 9))


def test_sintel_discovery_split_lists_and_items(tmp_path):
    root = str(tmp_path / "sintel")
    truth = _sintel_tree(root)
    random.seed(3)
    tr = D.SintelClean(root, "train", crop_type="center", crop_shape=(32, 48))
    # 2 scenes x 3 consecutive pairs = 6 samples, 90 / 10 split, both lists persisted
    assert len(tr) == 5 and tr.image_size == (32, 48)
    lines = open(f"{root}/train.txt").read().splitlines() + open(f"{root}/val.txt").read().splitlines()
    assert len(lines) == 6 and len(set(lines)) == 6
    for ln in lines:
        a, b, f = ln.split(",")
        assert "/clean/" in a and os.path.dirname(a) == os.path.dirname(b)               # pairs never cross scenes
        assert int(b[-8:-4]) == int(a[-8:-4]) + 1
        assert f == a.replace("clean", "flow").replace(".png", ".flo")
    va = D.SintelClean(root, "val", crop_type="center", crop_shape=(32, 48))              # second construction reads the lists
    assert len(va) == 1 and set(tr.samples).isdisjoint(va.samples)
    fin = D.SintelFinal(root, "train", crop_type="center", crop_shape=(32, 48))           # same lists, mapped to the final pass
    assert all("/final/" in s[0] and "/final/" in s[1] and "/flow/" in s[2] for s in fin.samples)
    images, flow = tr[0]
    a, b, f = tr.samples[0]
    assert images.shape == (2, 32, 48, 3) and images.dtype == np.uint8 and flow.shape == (32, 48, 2) and flow.dtype == np.float32
    np.testing.assert_array_equal(images[0], truth[a][4:36, 4:52])
    np.testing.assert_array_equal(images[1], truth[b][4:36, 4:52])
    np.testing.assert_array_equal(flow, truth[f][4:36, 4:52])
    # random crops: one offset pair per item, shared by both images and the flow
    rnd = D.SintelClean(root, "train", crop_type="random", crop_shape=(16, 24))
    random.seed(11)
    images, flow = rnd[1]
    random.seed(11)
    h1, w1 = random.randint(0, 40 - 16), random.randint(0, 56 - 24)
    a, b, f = rnd.samples[1]
    np.testing.assert_array_equal(images[0], truth[a][h1:h1 + 16, w1:w1 + 24])
    np.testing.assert_array_equal(flow, truth[f][h1:h1 + 16, w1:w1 + 24])
    # no crop: full frames; the resize_shape argument only enters image_size (flow.py:76)
    full = D.SintelClean(root, "train", crop_type=None, crop_shape=None, resize_shape=(20, 28))
    images, flow = full[0]
    assert full.image_size == (20, 28) and images.shape == (2, 40, 56, 3)
    # through a torch DataLoader as train.py:36-41 builds it
    from torch.utils import data
    batch_images, batch_flows = next(iter(data.DataLoader(tr, batch_size=2, shuffle=False, drop_last=True)))
    assert tuple(batch_images.shape) == (2, 2, 32, 48, 3) and str(batch_images.dtype) == "torch.uint8"
    assert tuple(batch_flows.shape) == (2, 32, 48, 2)


def test_flying_chairs_and_kitti(tmp_path):
    rng = np.random.default_rng(1)
    root = str(tmp_path / "chairs")
    os.makedirs(root + "/data")
    for i in range(1, 4):
        for j in (1, 2):
            cv2.imwrite(f"{root}/data/{i:05d}_img{j}.ppm", rng.integers(0, 256, (24, 32, 3), dtype=np.uint8))
        save_flow(f"{root}/data/{i:05d}_flow.flo", rng.normal(0, 2, (24, 32, 2)).astype(np.float32))
    random.seed(0)
    ch = D.get_dataset("FlyingChairs")(root, "train", crop_type="center", crop_shape=(16, 16))
    assert len(ch) == 2
    for a, b, f in ch.samples:
        assert a.endswith("_img1.ppm") and b == a.replace("img1", "img2") and f == a.replace("img1", "flow").replace(".ppm", ".flo")
    images, flow = ch[0]
    assert images.shape == (2, 16, 16, 3) and flow.shape == (16, 16, 2)
    # KITTI: 16-bit PNG, BGR = (valid, v, u); (x - 2^15) / 64
    kroot = str(tmp_path / "kitti")
    os.makedirs(kroot + "/training/image_2"); os.makedirs(kroot + "/training/flow_occ")
    u = rng.uniform(-30, 30, (12, 20)); v = rng.uniform(-10, 10, (12, 20))
    valid = rng.random((12, 20)) > 0.3
    u[0, 0], v[0, 0], valid[0, 0] = 0.0, 0.0, True                                       # exact zero -> 1e-10 (flow.py:257)
    enc = np.zeros((12, 20, 3), np.uint16)
    enc[..., 2] = np.round(u * 64 + 2 ** 15).astype(np.uint16)
    enc[..., 1] = np.round(v * 64 + 2 ** 15).astype(np.uint16)
    enc[..., 0] = valid.astype(np.uint16)
    cv2.imwrite(kroot + "/training/flow_occ/000000_10.png", enc)
    fl = D.load_kitti_flow(kroot + "/training/flow_occ/000000_10.png")
    assert fl.dtype == np.float32 and fl.shape == (12, 20, 2)
    np.testing.assert_allclose(fl[..., 0][valid], (np.round(u * 64) / 64)[valid], atol=1e-6)
    np.testing.assert_allclose(fl[..., 1][valid], (np.round(v * 64) / 64)[valid], atol=1e-6)
    assert (fl[~valid] == 0).all() and fl[0, 0, 0] == np.float32(1e-10)
    kd = D.KITTI.__new__(D.KITTI)
    kd.dataset_dir, kd.train_or_val = kroot, "train"
    random.seed(2)
    kd.has_no_txt()
    assert len(kd.samples) == 180 and all(s[2].endswith("_10.png") and "/flow_occ/" in s[2] for s in kd.samples)
    assert sorted(D.get_dataset(n).__name__ for n in ("FlyingChairs", "Sintel", "SintelClean", "SintelFinal", "KITTI")) == \
        ["FlyingChairs", "KITTI", "Sintel", "SintelClean", "SintelFinal"]


def test_experiment_saver_and_summary_writer(tmp_path, monkeypatch):
    import argparse
    import json
    from pwcnet_b200.utils import ExperimentSaver, SummaryWriter, save_config
    monkeypatch.chdir(tmp_path)
    os.makedirs("model"); open("model/model_1.ckpt.index", "w").write("x")
    sv = ExperimentSaver(logdir=str(tmp_path / "logs" / "history_x"), parse_args=argparse.Namespace(lr=1e-4, dataset="SintelClean"))
    sv.append(["./figure", "./model"])
    sv.save()
    assert json.load(open(tmp_path / "logs/history_x/config.json")) == {"lr": 1e-4, "dataset": "SintelClean"}
    assert (tmp_path / "logs/history_x/model/model_1.ckpt.index").exists() and not os.path.exists("model")
    w = SummaryWriter(str(tmp_path / "logs/history_x/train"))
    w.add_summary({"loss/pwc": 1.5, "EPE/source": 0.25}, 1000)
    w.close()
    rec = json.loads(open(tmp_path / "logs/history_x/train/scalars.jsonl").read())
    assert rec["step"] == 1000 and rec["loss/pwc"] == 1.5 and rec["EPE/source"] == 0.25
    with pytest.raises(TypeError):
        save_config([1, 2])
