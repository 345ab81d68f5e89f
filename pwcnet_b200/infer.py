"""Inference entry point: the reference's test.py (Tester, test.py:19-65) on the B200 compute path.

    python -m pwcnet_b200.infer --input_images a.png b.png [--resume model_250.ckpt] [--time] [--out flow.flo]

Same flow as the reference: read two images, `factor_crop` to multiples of 64, scale to [0,1], run PWCDCNet,
rescale pyramid level l by 20/2**(num_levels-l) (test.py:57-60).  Instead of the matplotlib figure (out of scope,
SURVEY 8) the final flow is written as a Middlebury `.flo` file; `--time` reproduces the reference's timing loop
(test.py:47-53) with CUDA events."""
from __future__ import annotations

import argparse
import time

import numpy as np
import torch

from .flow_io import factor_crop, save_flow
from .model import PWCDCNet


def _imread(path):
    import cv2   # only the CLI needs an image decoder
    img = cv2.imread(path, cv2.IMREAD_COLOR)
    if img is None:
        raise FileNotFoundError(path)
    return cv2.cvtColor(img, cv2.COLOR_BGR2RGB)


class Tester(object):
    def __init__(self, args):
        self.args = args
        img1, img2 = (factor_crop(_imread(p)) for p in args.input_images)
        if img1.shape != img2.shape:
            raise ValueError(f"input images differ in shape: {img1.shape} vs {img2.shape}")
        # uint8 (2, h, w, 3): the `/255.0` of test.py:32 runs on the device (bit-identical, see PWCDCNet.__call__)
        self.images = np.ascontiguousarray(np.array([img1, img2]))
        self.model = PWCDCNet()
        if args.resume is not None:
            print(f'Loading learned model from checkpoint {args.resume}')
            self.model.load_weights(args.resume)
        else:
            print('!!! Test with un-learned model !!!')

    def test(self):
        im0, im1 = self.images[0:1], self.images[1:2]
        flow_final, flows = self.model(im0, im1)
        if self.args.time:
            n = self.args.iters
            torch.cuda.synchronize()
            t0 = time.time()
            for _ in range(n):
                flow_final, flows = self.model(im0, im1)
            torch.cuda.synchronize()
            print(f'Inference time: {(time.time() - t0) / n} sec (averaged over {n} iterations)')
        flow_set = [f[0].cpu().numpy() * (20 / 2 ** (self.model.num_levels - l)) for l, f in enumerate(flows)]
        out = self.args.out
        save_flow(out, flow_final[0].cpu().numpy())
        print(f'Flow saved to {out} (pyramid shapes {[f.shape for f in flow_set]})')
        return flow_final, flow_set


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--input_images', type=str, nargs=2, required=True, help='Target images (required)')
    parser.add_argument('--resume', type=str, default=None, help='Learned parameter checkpoint file [None]')
    parser.add_argument('--time', '-t', action='store_true', help='Stored option for inference speed measurement')
    parser.add_argument('--iters', type=int, default=1000)
    parser.add_argument('--out', type=str, default='flow.flo')
    args = parser.parse_args(argv)
    for key, item in vars(args).items():
        print(f'{key} : {item}')
    Tester(args).test()


if __name__ == '__main__':
    main()
