"""Drop-in equivalents of the reference's `modules.py` callables on the B200 compute path.

Same class names, constructor arguments, call signatures and return structure as the reference
(daigo0927/pwcnet modules.py); tensors are NHWC float32 torch CUDA tensors instead of TF graph
tensors, and variables live in a `params` dict keyed by the reference's checkpoint names
(`<scope>/<name>/conv2d[_i]/{kernel,bias}`) instead of TF variable scopes.

These stand-alone modules allocate fresh outputs per call (like TF ops do); `PWCDCNet.__call__`
in model.py runs the same kernels through a pre-planned workspace with concat buffers and the
fused warp+cost-volume kernel.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import ops

PYRAMID_FILTERS = [16, 32, 64, 96, 128, 192]      # modules.py:45
ESTIMATOR_FILTERS = [128, 128, 96, 64, 32]        # modules.py:234
CONTEXT_FILTERS = [128, 128, 128, 96, 64, 32, 2]  # modules.py:306-325
CONTEXT_DILATIONS = [1, 2, 4, 8, 16, 1, 1]


def _layer(params: Dict[str, torch.Tensor], scope: str, idx: int):
    name = f"{scope}/conv2d" + (f"_{idx}" if idx else "")
    try:
        return params[name + "/kernel"], params[name + "/bias"]
    except KeyError as e:
        raise KeyError(f"missing variable {e} (expected the reference's checkpoint naming)") from None


def _conv_block(filters, kernel_size=(3, 3), strides=(1, 1), batch_norm=False, *, kernel=None, bias=None):
    """modules.py:7-15: Conv2D(filters, 3x3, strides, 'same') [+ BatchNormalization] + leaky_relu(0.2).  TF creates the
    variables when the closure runs; here they are passed in (`kernel` HWIO with `filters` output channels, `bias`).
    batch_norm=True is not supported (no caller in the reference sets it: modules.py:210 fixes it to False)."""
    if tuple(kernel_size) != (3, 3):
        raise NotImplementedError("_conv_block: only 3x3 kernels exist in the reference")
    if batch_norm:
        raise NotImplementedError("_conv_block: batch_norm=True has no caller in the reference (modules.py:210)")

    def f(x):
        if kernel is None or bias is None or kernel.shape[3] != filters:
            raise ValueError("_conv_block: pass kernel (3,3,Cin,filters) and bias (filters,)")
        return ops.conv3x3(x, kernel, bias, stride=strides[0], alpha=0.2)
    return f


class FeaturePyramidExtractor(object):
    """Feature pyramid extractor module simple/original (modules.py:19-39): two convs per level."""

    def __init__(self, num_levels=6, name='fp_extractor', params=None, scope='pwcnet'):
        self.num_levels = num_levels
        self.filters = list(PYRAMID_FILTERS)
        self.name = name
        self.params = params
        self.scope = f"{scope}/{name}"

    def __call__(self, x, reuse=True):
        feature_pyramid = []
        for l in range(self.num_levels):
            for j, stride in enumerate((2, 1)):
                k, b = _layer(self.params, self.scope, 2 * l + j)
                x = ops.conv3x3(x, k, b, stride=stride, alpha=0.1)
            feature_pyramid.append(x)
        return feature_pyramid[::-1]


class OpticalFlowEstimator(object):
    """Optical flow estimator module simple/original (modules.py:208-224): concat [cost, x, flow] -> five
    _conv_block (leaky 0.2) -> 2-channel conv; returns (feature, flow)."""

    def __init__(self, name='of_estimator', params=None, scope='pwcnet'):
        self.batch_norm = False
        self.name = name
        self.params = params
        self.scope = f"{scope}/{name}"

    def __call__(self, cost, x, flow):
        x = torch.cat([cost, x, flow.to(torch.float32)], dim=3)
        for i, f in enumerate(ESTIMATOR_FILTERS):
            k, b = _layer(self.params, self.scope, i)
            x = _conv_block(f, (3, 3), (1, 1), self.batch_norm, kernel=k, bias=b)(x)
        feature = x
        k, b = _layer(self.params, self.scope, len(ESTIMATOR_FILTERS))
        flow = ops.conv3x3(feature, k, b, alpha=1.0)
        return feature, flow


class FeaturePyramidExtractor_custom(object):
    """Feature pyramid extractor module (modules.py:42-71)."""

    def __init__(self, num_levels=6, name='fp_extractor', params=None, scope='pwcdcnet'):
        self.num_levels = num_levels
        self.filters = list(PYRAMID_FILTERS)
        self.name = name
        self.params = params
        self.scope = f"{scope}/{name}"

    def __call__(self, images, reuse=True):
        """images (batch,h,w,3) -> features_pyramid, deep -> shallow order."""
        features_pyramid = []
        x = images
        for l in range(self.num_levels):
            for j, stride in enumerate((2, 1, 1)):
                k, b = _layer(self.params, self.scope, 3 * l + j)
                x = ops.conv3x3(x, k, b, stride=stride, alpha=0.1)
            features_pyramid.append(x)
        return features_pyramid[::-1]


def nearest_warp(x, flow):
    """modules.py:83-97."""
    return ops.warp(x, flow, 1.0, 'nearest')


def bilinear_warp(x, flow):
    """modules.py:99-137."""
    return ops.warp(x, flow, 1.0, 'bilinear')


class WarpingLayer(object):
    """modules.py:139-154."""

    def __init__(self, warp_type='nearest', name='warping'):
        self.warp = warp_type
        self.name = name

    def __call__(self, x, flow):
        assert self.warp in ['nearest', 'bilinear']
        return nearest_warp(x, flow) if self.warp == 'nearest' else bilinear_warp(x, flow)


class CostVolumeLayer(object):
    """Cost volume module (modules.py:183-204)."""

    def __init__(self, search_range=4, name='cost_volume'):
        self.s_range = search_range
        self.name = name

    def __call__(self, features_0, features_0from1):
        return ops.cost_volume(features_0, features_0from1, self.s_range, alpha=0.1)


class OpticalFlowEstimator_custom(object):
    """Optical flow estimator module (modules.py:227-285)."""

    def __init__(self, use_dc=False, name='of_estimator', params=None, scope='pwcdcnet'):
        self.filters = list(ESTIMATOR_FILTERS)
        self.use_dc = use_dc
        self.name = name
        self.params = params
        self.scope = f"{scope}/{name}"

    def __call__(self, cv, features_0=None, flows_up_prev=None, features_up_prev=None, is_output=False):
        parts = [cv] + [f for f in (features_0, flows_up_prev, features_up_prev) if f is not None]
        features = torch.cat(parts, dim=3) if len(parts) > 1 else cv
        for i, _ in enumerate(self.filters):
            k, b = _layer(self.params, self.scope, i)
            conv = ops.conv3x3(features, k, b, alpha=0.1)
            features = torch.cat([conv, features], dim=3) if self.use_dc else conv
        k, b = _layer(self.params, self.scope, len(self.filters))
        flows = ops.conv3x3(features, k, b, alpha=1.0, residual=flows_up_prev)   # residual: modules.py:275-277
        if is_output:
            return flows, features
        _, h, w, _ = flows.shape
        flows_up = ops.resize_bilinear(flows, 2 * h, 2 * w)
        features_up = ops.resize_bilinear(features, 2 * h, 2 * w)
        return flows, flows_up, features_up


class ContextNetwork(object):
    """Context module (modules.py:290-326)."""

    def __init__(self, name='context', params=None, scope='pwcdcnet'):
        self.name = name
        self.params = params
        self.scope = f"{scope}/{name}"

    def __call__(self, flows, features):
        x = torch.cat([flows, features], dim=3)
        n = len(CONTEXT_FILTERS)
        for i, d in enumerate(CONTEXT_DILATIONS):
            k, b = _layer(self.params, self.scope, i)
            last = i == n - 1
            x = ops.conv3x3(x, k, b, dilation=d, alpha=1.0 if last else 0.1, residual=flows if last else None)
        return x
