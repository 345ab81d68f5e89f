"""CPU check of `Trainer._grad_buffers`: the gradient workspace mirrors the plan's activation buffers one to one, the
buffers whose first writer accumulates sit in front (one fill clears exactly them), nothing overlaps, every view starts on
a 256-byte boundary.  (That the OTHER buffers are overwritten completely is a GPU test: test_gpu_train.py.)"""
import types

import torch

from pwcnet_b200.train import Trainer


def _fake_plan(use_dc):
    t = lambda *s: torch.empty(s)
    p = types.SimpleNamespace(B=1, H=64, W=128)
    p.pyr = [[t(2, 32 >> l, 64 >> l, c) for _ in range(3)] for l, c in enumerate((16, 32, 64))]
    levels = 2
    p.S = [t(1, 8 << l, 16 << l, 100 + 7 * l) for l in range(levels)]
    p.tmp = [None if use_dc else [t(1, 8 << l, 16 << l, c) for c in (128, 128, 96, 64, 36)] for l in range(levels)]
    p.flows = [t(1, 8 << l, 16 << l, 2) for l in range(levels)]
    p.f1w = [None] + [t(1, 8 << l, 16 << l, 32) for l in range(1, levels)]
    p.ctx = [t(1, 16, 32, c) for c in (128, 128, 128, 96, 64, 32)]
    return p


def _check(use_dc):
    tr = object.__new__(Trainer)
    tr._gbufs = {}
    tr.model = types.SimpleNamespace(device=torch.device("cpu"))
    p = _fake_plan(use_dc)
    g = tr._grad_buffers(p)
    assert tr._grad_buffers(p) is g                      # cached per shape
    base = g.flat.data_ptr()
    spans = []

    def add(view, act, cleared):
        assert view.shape == act.shape and view.is_contiguous()
        off = (view.data_ptr() - base) // 4
        assert off % 64 == 0                             # 256-byte aligned
        assert (off + view.numel() <= g.n_clear) if cleared else (off >= g.n_clear)
        spans.append((off, off + view.numel()))

    for lev, glev in zip(p.pyr, g.pyr):
        for j, (a, v) in enumerate(zip(lev, glev)):
            add(v, a, cleared=(j == 2))                  # only the level's output collects several gradients
    for l in range(len(p.S)):
        add(g.S[l], p.S[l], True)
        tmp = p.tmp[l] or []
        assert len(g.tmp[l]) == len(tmp)
        for j, (a, v) in enumerate(zip(tmp, g.tmp[l])):
            add(v, a, cleared=(j == len(tmp) - 1))       # [features | flows | pad]: head, context conv 0 and the residual accumulate
        add(g.flows[l], p.flows[l], True)
        if p.f1w[l] is None:
            assert g.f1w[l] is None
        else:
            add(g.f1w[l], p.f1w[l], True)
    for a, v in zip(p.ctx, g.ctx):
        add(v, a, False)
    spans.sort()
    assert all(a1 <= b0 for (_, a1), (b0, _) in zip(spans, spans[1:]))   # no overlap
    assert spans[-1][1] <= g.flat.numel() and 0 < g.n_clear < g.flat.numel()
    assert not g.flat.any()                              # allocated zeroed: the first step needs no special case


def test_grad_buffer_layout_plain_stack():
    _check(use_dc=False)


def test_grad_buffer_layout_dense_stack():
    _check(use_dc=True)
