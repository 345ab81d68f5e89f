"""Drop-in equivalents of the reference's `losses.py` on CUDA tensors (NHWC float32).
Each function returns a 0-dim CUDA tensor (one fused reduction kernel per pyramid level)."""
from __future__ import annotations

import torch

from . import ops


def _acc(like):
    return torch.zeros((), dtype=torch.float32, device=like.device)


def L1loss(x, y):
    """losses.py:4-5 : mean_b sum_{h,w} ||x - y||_1 (x and y of the same shape)."""
    return ops.lploss_level(x.contiguous(), y, 1.0, _acc(x), gt_div=1.0, ord=1)


def L2loss(x, y):
    """losses.py:7-8 : mean_b sum_{h,w} ||x - y||_2."""
    return ops.lploss_level(x.contiguous(), y, 1.0, _acc(x), gt_div=1.0, ord=2)


def EPE(flows_gt, flows):
    """losses.py:11-13 : mean over (b,h,w) of ||flows_gt - flows||_2 (both unscaled)."""
    return ops.epe(flows_gt.contiguous(), flows.contiguous(), _acc(flows_gt))


def multiscale_loss(flows_gt, flows_pyramid, weights, name='multiscale_loss'):
    """losses.py:15-31 : sum_l w_l * L2loss(resize_nearest(flows_gt/20, (h_l,w_l)), flows_pyramid[l])."""
    acc = _acc(flows_gt)
    gt = flows_gt.contiguous()
    for weight, fs in zip(weights, flows_pyramid):
        ops.lploss_level(gt, fs, weight, acc, gt_div=20.0, ord=2)
    return acc


def multirobust_loss(flows_gt, flows_pyramid, weights, epsilon=0.01, q=0.4, name='multirobust_loss'):
    """losses.py:33-47.  The reference raises NameError here (`loss_level` is undefined, losses.py:45);
    the evident intent, sum_l w_l * (L1loss_l + epsilon)**q, is what this computes."""
    gt = flows_gt.contiguous()
    loss = _acc(flows_gt)
    for weight, fs in zip(weights, flows_pyramid):
        lvl = ops.lploss_level(gt, fs, 1.0, _acc(flows_gt), gt_div=20.0, ord=1)
        loss = loss + weight * (lvl + epsilon) ** q
    return loss
