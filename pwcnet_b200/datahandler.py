"""Optical-flow datasets for training / evaluation: the data side of the reference (its `datahandler` submodule,
datahandler/flow.py + datahandler/utils.py) restated for this package.

Same class names, constructor arguments and item format as the reference, so `train.py`'s
`get_dataset(name)(train_or_val=..., dataset_dir=..., crop_type=..., crop_shape=..., resize_shape=..., resize_scale=...)`
+ `torch.utils.data.DataLoader` (train.py:28-41) works unchanged:

    item = (images uint8 (2, h, w, 3) RGB, flow float32 (h, w, 2) in pixels, channel 0 = x)

The uint8 images go to the device as bytes (`Trainer.step` / `TrainStream` / `PWCDCNet.__call__` accept uint8 and apply the
reference's `/255.0`, train.py:122, on the device).  Image decoding uses OpenCV (the reference uses imageio; PNG / PPM
decode to the same bytes).  Behaviours kept on purpose, each with its reference line:
  * sample lists come from `<dataset_dir>/train.txt` / `val.txt` (`img0,img1,flow` per line) when present; otherwise the
    dataset is discovered, shuffled with the global `random` module, split 90 / 10 and the two lists are written
    (flow.py:72-74,118-130);
  * the crop offsets are drawn once per sample with `random.randint` (rows first) and applied to both images and the flow
    (utils.py:6-14, flow.py:88-91); the centre crop uses floor division on both ends (utils.py:17-24);
  * `self.resize_shape = crop_shape` (flow.py:76): the `resize_shape` argument only enters `image_size`; items are resized to
    the CROP shape (an identity when a crop was taken, nothing when `crop_shape` is None);
  * flow resizing scales u by the width ratio and v by the height ratio (flow.py:29-37); `rescale_flow` multiplies
    (u, v) by `resize_scale` in the order given (flow.py:39-47);
  * KITTI flow PNGs: 16-bit BGR, u = (R - 2^15) / 64, v = (G - 2^15) / 64, |value| < 1e-10 -> 1e-10, B == 0 -> invalid -> 0
    (flow.py:251-259).
"""
from __future__ import annotations

import os
import random
from itertools import groupby, islice
from pathlib import Path

import numpy as np

try:  # the dataset classes are torch Datasets when torch is importable (it always is in this package)
    from torch.utils.data import Dataset
except Exception:  # pragma: no cover
    Dataset = object

from .flow_io import load_flow as _load_flo


# ---------------------------------------------------------------------------------------------- utils.py
class StaticRandomCrop(object):
    """utils.py:6-14: offsets fixed at construction so that several arrays get the same crop."""

    def __init__(self, image_size, crop_size):
        self.th, self.tw = crop_size
        h, w = image_size
        self.h1 = random.randint(0, h - self.th)
        self.w1 = random.randint(0, w - self.tw)

    def __call__(self, image):
        return image[self.h1:self.h1 + self.th, self.w1:self.w1 + self.tw]


class StaticCenterCrop(object):
    """utils.py:17-24."""

    def __init__(self, image_size, crop_size):
        self.th, self.tw = crop_size
        self.h, self.w = image_size

    def __call__(self, image):
        top, left = (self.h - self.th) // 2, (self.w - self.tw) // 2
        return image[top:(self.h + self.th) // 2, left:(self.w + self.tw) // 2]


def window(seq, n=2):
    """utils.py:27-36: sliding windows of width n: (s0..s[n-1]), (s1..sn), ..."""
    it = iter(seq)
    cur = tuple(islice(it, n))
    if len(cur) == n:
        yield cur
    for elem in it:
        cur = cur[1:] + (elem,)
        yield cur


def get_size(origin_size=None, crop_size=None, resize_size=None, resize_scale=None):
    """utils.py:43-58: the item size after cropping / resizing; priority resize > crop > origin."""
    image_size = resize_size if resize_size is not None else crop_size if crop_size is not None else (origin_size or None)
    if image_size is None:
        raise ValueError('One of the argument should be not None')
    if resize_scale is not None:
        image_size = (image_size[0] * resize_scale[0], image_size[1] * resize_scale[1])
    return image_size


# ---------------------------------------------------------------------------------------------- flow.py helpers
def load_flow(uri):
    """flow.py:15-28 (`.flo`, Middlebury): (h, w, 2) float32 or None on a bad magic number."""
    return _load_flo(str(uri))


def _cv2():
    import cv2
    return cv2


def _imread_rgb(path):
    cv2 = _cv2()
    img = cv2.imread(str(path), cv2.IMREAD_COLOR)
    if img is None:
        raise FileNotFoundError(str(path))
    return cv2.cvtColor(img, cv2.COLOR_BGR2RGB)


def resize_flow(flow, resize_shape):
    """flow.py:29-37: bilinear resize to (th, tw); u scaled by tw / w, v by th / h."""
    if flow.ndim != 3:
        raise ValueError(f'Flow dimension should be 3, but found {flow.ndim} dimension')
    h, w = flow.shape[:2]
    th, tw = resize_shape
    ratio = np.array([tw / w, th / h]).reshape((1, 1, 2))
    return np.float32(_cv2().resize(flow, dsize=(tw, th)) * ratio)


def rescale_flow(flow, resize_scale):
    """flow.py:39-47: target (int(h s0), int(w s1)); (u, v) multiplied by (s0, s1) as the reference does."""
    if flow.ndim != 3:
        raise ValueError(f'Flow dimension should be 3, but found {flow.ndim} dimension')
    h, w = flow.shape[:2]
    th, tw = int(h * resize_scale[0]), int(w * resize_scale[1])
    ratio = np.array(resize_scale).reshape((1, 1, 2))
    return np.float32(_cv2().resize(flow, dsize=(tw, th)) * ratio)


def load_kitti_flow(uri):
    """flow.py:251-259."""
    cv2 = _cv2()
    raw = cv2.imread(str(uri), cv2.IMREAD_UNCHANGED)
    if raw is None:
        raise FileNotFoundError(str(uri))
    flow = raw[:, :, 2:0:-1].astype(np.float32)          # (R, G) = (u, v)
    invalid = raw[:, :, 0] == 0
    flow = (flow - 2 ** 15) / 64
    flow[np.abs(flow) < 1e-10] = 1e-10
    flow[invalid] = 0
    return flow


# ---------------------------------------------------------------------------------------------- datasets
class BaseDataset(Dataset):
    """flow.py:51-130."""

    def __init__(self, dataset_dir, train_or_val, origin_size=None, crop_type='random', crop_shape=None,
                 resize_shape=None, resize_scale=None):
        self.dataset_dir = dataset_dir
        assert train_or_val in ['train', 'val'], 'Argument should be either of [train, val]'
        self.train_or_val = train_or_val
        self.image_size = get_size(origin_size, crop_shape, resize_shape, resize_scale)
        self.crop_type = crop_type
        self.crop_shape = crop_shape
        self.resize_shape = crop_shape                   # sic, flow.py:76
        self.resize_scale = resize_scale
        if (Path(dataset_dir) / (train_or_val + '.txt')).exists():
            self.has_txt()
        else:
            self.has_no_txt()

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, idx):
        img0_path, img1_path, flow_path = self.samples[idx]
        image_0, image_1 = _imread_rgb(img0_path), _imread_rgb(img1_path)
        flow = self.load_flow(flow_path)
        if self.crop_shape is not None:
            crop_cls = StaticRandomCrop if self.crop_type == 'random' else StaticCenterCrop
            cropper = crop_cls(image_0.shape[:2], self.crop_shape)
            image_0, image_1, flow = cropper(image_0), cropper(image_1), cropper(flow)
        if self.resize_shape is not None and tuple(image_0.shape[:2]) != tuple(self.resize_shape):
            cv2 = _cv2()
            size = tuple(self.resize_shape[::-1])
            image_0, image_1 = cv2.resize(image_0, dsize=size), cv2.resize(image_1, dsize=size)
            flow = resize_flow(flow, self.resize_shape)
        if self.resize_scale is not None:
            cv2 = _cv2()
            sx, sy = self.resize_scale
            image_0 = cv2.resize(image_0, dsize=(0, 0), fx=sx, fy=sy)
            image_1 = cv2.resize(image_1, dsize=(0, 0), fx=sx, fy=sy)
            flow = rescale_flow(flow, self.resize_scale)
        images = np.stack([np.ascontiguousarray(image_0), np.ascontiguousarray(image_1)], axis=0)
        return images, np.ascontiguousarray(flow, dtype=np.float32)

    def _read_list(self, fix=lambda p: p):
        self.samples = []
        with open(Path(self.dataset_dir) / (self.train_or_val + '.txt'), 'r') as f:
            for line in f.readlines():
                if not line.strip():
                    continue
                img_0_path, img_1_path, flow_path = line.split(',')
                self.samples.append((fix(img_0_path), fix(img_1_path), flow_path.strip()))

    def has_txt(self):
        self._read_list()

    def has_no_txt(self):
        raise NotImplementedError

    def split(self, samples, val_ratio=0.1):
        """flow.py:118-130: shuffle (global `random`), 90 / 10, persist both lists."""
        root = Path(self.dataset_dir)
        random.shuffle(samples)
        cut = int(len(samples) * (1 - val_ratio))
        parts = {'train': samples[:cut], 'val': samples[cut:]}
        for name, part in parts.items():
            with open(root / (name + '.txt'), 'w') as f:
                f.writelines(','.join(s) + '\n' for s in part)
        self.samples = parts[self.train_or_val]

    def load_flow(self, flow_path):
        return load_flow(flow_path)


class FlyingChairs(BaseDataset):
    """flow.py:135-148: `<dir>/data/NNNNN_img1.ppm, NNNNN_img2.ppm, NNNNN_flow.flo`."""

    def has_no_txt(self):
        imgs = sorted((Path(self.dataset_dir) / 'data').glob('*.ppm'))
        samples = [(str(a), str(b), str(a).replace('img1', 'flow').replace('.ppm', '.flo'))
                   for a, b in zip(imgs[::2], imgs[1::2])]
        self.split(samples)


class FlyingThings3D(BaseDataset):
    """flow.py:153-162: a stub in the reference as well (its has_no_txt is `pass`: only list files work)."""

    def has_no_txt(self):
        self.samples = []


class Sintel(BaseDataset):
    """flow.py:167-186: `<dir>/training/<mode>/<scene>/frame_NNNN.png`, flow in `<dir>/training/flow/<scene>/frame_NNNN.flo`;
    consecutive frames of one scene form a pair."""

    def __init__(self, dataset_dir, train_or_val, mode='clean', origin_size=None, crop_type='random', crop_shape=None,
                 resize_shape=None, resize_scale=None):
        self.mode = mode
        super().__init__(dataset_dir, train_or_val, origin_size, crop_type, crop_shape, resize_shape, resize_scale)

    def has_no_txt(self):
        frames = sorted(map(str, (Path(self.dataset_dir) / 'training' / self.mode).glob('**/*.png')))
        samples = []
        for _, scene in groupby(frames, lambda s: s.split('/')[-2]):
            for a, b in window(list(scene), 2):
                samples.append((a, b, a.replace(self.mode, 'flow').replace('.png', '.flo')))
        self.split(samples)


class SintelClean(Sintel):
    """flow.py:188-205: list files written by either pass are usable: image paths are mapped to the clean pass."""

    def __init__(self, dataset_dir, train_or_val, origin_size=None, crop_type='random', crop_shape=None,
                 resize_shape=None, resize_scale=None):
        super().__init__(dataset_dir, train_or_val, 'clean', origin_size, crop_type, crop_shape, resize_shape, resize_scale)

    def has_txt(self):
        self._read_list(lambda p: p.replace('final', 'clean'))


class SintelFinal(Sintel):
    """flow.py:207-224."""

    def __init__(self, dataset_dir, train_or_val, origin_size=None, crop_type='random', crop_shape=None,
                 resize_shape=None, resize_scale=None):
        super().__init__(dataset_dir, train_or_val, 'final', origin_size, crop_type, crop_shape, resize_shape, resize_scale)

    def has_txt(self):
        self._read_list(lambda p: p.replace('clean', 'final'))


class KITTI(BaseDataset):
    """flow.py:229-259: KITTI 2015, `training/image_2/NNNNNN_10.png`, `_11.png`, flow `training/flow_occ/NNNNNN_10.png`."""

    N_PAIRS = 200

    def has_no_txt(self):
        root = Path(self.dataset_dir)
        samples = []
        for i in range(self.N_PAIRS):
            stem = str(i).zfill(6)
            samples.append((str(root / 'training/image_2' / f'{stem}_10.png'), str(root / 'training/image_2' / f'{stem}_11.png'),
                            str(root / 'training/flow_occ' / f'{stem}_10.png')))
        self.split(samples)

    def load_flow(self, uri):
        return load_kitti_flow(uri)


def get_dataset(dataset_name):
    """flow.py:262-278."""
    return {"FlyingChairs": FlyingChairs, "Sintel": Sintel, "SintelClean": SintelClean, "SintelFinal": SintelFinal,
            "KITTI": KITTI}[dataset_name]
