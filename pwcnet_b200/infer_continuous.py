"""Continuous inference over an image sequence: the reference's test_continuous.py (Tester, test_continuous.py:19-66) on
the B200 compute path.

    python -m pwcnet_b200.infer_continuous -i frame_0001.png frame_0002.png frame_0003.png ... [-r model_250.ckpt]
                                           [--out_dir ./test_figure]

Same flow as the reference: the images are taken as a sliding window of consecutive pairs (test_continuous.py:46),
every image is `factor_crop`ped to multiples of 64 (the reference's placeholder is (1, 2, None, None, 3), so sizes may
change along the sequence: a workspace + CUDA graph is planned per distinct size), the network runs on each pair, and
pyramid level l is rescaled by 20/2**(num_levels-l) (test_continuous.py:54-57).  Instead of the matplotlib figure
`./test_figure/<dname>/<fname>.png` (out of scope) the full-resolution flow of every pair is written as
`<out_dir>/<dname>/<fname>.flo`.  Images cross PCIe as uint8; the `/255.` of test_continuous.py:49 runs on the device.
Pairs of equal size are pipelined through `InferenceStream` (H2D of pair i+1 and D2H of pair i-1 overlap pair i)."""
from __future__ import annotations

import argparse
import os
import re

import numpy as np

from .flow_io import factor_crop, save_flow
from .infer import _imread
from .model import PWCDCNet
from .pipeline import InferenceStream


class Tester(object):
    def __init__(self, args):
        self.args = args
        self.model = PWCDCNet()
        if args.resume is not None:
            print(f'Loading learned model from checkpoint {args.resume}')
            self.model.load_weights(args.resume)
        else:
            print('!!! Test with un-learned model !!!')
        self.stream = InferenceStream(self.model, depth=2)

    def _write(self, img1_path, flow_final, flows):
        flow_set = [f[0].numpy() * (20 / 2 ** (self.model.num_levels - l)) for l, f in enumerate(flows)]
        parts = re.split('[/.]', img1_path)[-3:-1]                      # test_continuous.py:59
        dname, fname = parts if len(parts) == 2 else ('.', parts[-1])
        os.makedirs(os.path.join(self.args.out_dir, dname), exist_ok=True)
        out = os.path.join(self.args.out_dir, dname, fname + '.flo')
        save_flow(out, flow_final[0].numpy())
        return out, flow_set

    def test(self):
        os.makedirs(self.args.out_dir, exist_ok=True)
        paths = self.args.input_images
        written, pending = [], []
        prev = factor_crop(_imread(paths[0]))
        for img1_path, img2_path in zip(paths[:-1], paths[1:]):
            cur = factor_crop(_imread(img2_path))
            if cur.shape != prev.shape:
                raise ValueError(f"images of a pair differ in shape: {img1_path} {prev.shape} vs {img2_path} {cur.shape}")
            if pending and pending[-1][2] != prev.shape:                 # size changed: finish what is in flight
                while pending:
                    tk, pth, _ = pending.pop(0)
                    written.append(self._write(pth, *self.stream.collect(tk))[0])
            tk = self.stream.submit(np.ascontiguousarray(prev[None]), np.ascontiguousarray(cur[None]))
            pending.append((tk, img1_path, prev.shape))
            if len(pending) > 1:
                tk0, pth, _ = pending.pop(0)
                written.append(self._write(pth, *self.stream.collect(tk0))[0])
            prev = cur
        while pending:
            tk, pth, _ = pending.pop(0)
            written.append(self._write(pth, *self.stream.collect(tk))[0])
        print(f'{len(written)} flow fields saved under {self.args.out_dir}')
        return written


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('-i', '--input_images', type=str, nargs='+', required=True, help='Target images (required)')
    parser.add_argument('-r', '--resume', type=str, default=None, help='Learned parameter checkpoint file [None]')
    parser.add_argument('--out_dir', type=str, default='./test_figure')
    args = parser.parse_args(argv)
    if len(args.input_images) == 1 and '*' in args.input_images[0]:      # expand wild-card (test_continuous.py:76-78)
        from glob import glob
        args.input_images = sorted(glob(args.input_images[0]))
    if len(args.input_images) < 2:
        raise ValueError('# of input images must be >= 2')
    print(args.resume)
    for i, image in enumerate(args.input_images):
        print(image)
        if i == 5:
            print(f'... and more ({len(args.input_images)} images)')
            break
    return Tester(args).test()


if __name__ == '__main__':
    main()
