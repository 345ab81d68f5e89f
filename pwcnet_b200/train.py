"""Training step of the reference (train.py:65-92) on the B200 compute path.

    trainer = Trainer(model, lr=1e-4, gamma=4e-4, weights=[0.32, 0.08, 0.02, 0.01, 0.005])
    loss, loss_ms, epe = trainer.step(images_0, images_1, flows_gt)      # one sess.run([optimizer, loss, epe])

What the reference's graph does per step, and what runs here instead:
  * forward PWCDCNet (model.py:95-134)                    -> model's forward launches (activations stay in the plan)
  * loss = multiscale_loss + gamma * sum_v l2_loss(v) over all 110 variables incl. biases (train.py:66-75),
    epe = EPE(flows_gt, flows_final) (train.py:77)         -> fused loss kernels
  * tf.gradients (37k-node backward graph)    -> `Trainer.backward`: ~190 launches of the kernels in
    csrc/backward.cu over gradient buffers that mirror the activation buffers; no autograd
  * data parallel: one all-reduce (sum) of the flat fp32 gradient (20.1 MB) over NCCL per step
  * AdamOptimizer.minimize, lr = piecewise_constant(global_step, [200k,250k,300k,350k,4M], lr/2^i)
    (train.py:82-89), global_step += 1 afterwards (train.py:91-92) -> one Adam launch over the flat buffers
  * the derived kernels (internal channel order, packed tensor-core tiles) are refreshed in place

`use_dc=True` (train.py:203-207, modules.py:269-270) trains through the same kernels over the dense per-level buffers."""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import ops, ops_bwd
from ._abi import PwcError
from .model import PWCDCNet, on_device
from .modules import CONTEXT_DILATIONS, CONTEXT_FILTERS, ESTIMATOR_FILTERS

DEFAULT_LOSS_WEIGHTS = [0.32, 0.08, 0.02, 0.01, 0.005]          # train.py:220-222
LR_BOUNDARIES = [200000, 250000, 300000, 350000, 4000000]       # train.py:83


def piecewise_lr(step: int, lr: float = 1e-4, boundaries: Sequence[int] = LR_BOUNDARIES) -> float:
    """tf.train.piecewise_constant(global_step, boundaries, [lr/2^i]) (train.py:82-85): value i applies for
    boundaries[i-1] < step <= boundaries[i]."""
    for i, b in enumerate(boundaries):
        if step <= b:
            return lr / (2 ** i)
    return lr / (2 ** len(boundaries))


class _Grads:
    """Gradient buffers mirroring one forward plan (one flat allocation, zeroed once per step)."""
    pass


class Trainer(object):
    def __init__(self, model: PWCDCNet, lr: float = 1e-4, gamma: float = 4e-4,
                 weights: Sequence[float] = DEFAULT_LOSS_WEIGHTS, beta1: float = 0.9, beta2: float = 0.999,
                 eps: float = 1e-8, lr_boundaries: Sequence[int] = LR_BOUNDARIES, process_group=None,
                 global_step: int = 0, tc_dgrad: Optional[bool] = None,
                 tc_wgrad: Optional[bool] = None):
        if model.fuse_warp:
            raise PwcError("Trainer needs the warped features in memory: construct the model with fuse_warp=False")
        if getattr(model, "cv_split", False) or getattr(model, "split_act", False):
            # the backward pass reads the warped fp32 features and the fp32 activations that the inference pipeline (split
            # cost volume, split conv -> conv chains) never materialises: training runs the default pipeline; plans made for
            # inference are dropped
            model.cv_split = False
            model.split_act = False
            model._plans.clear()
        if model.precision == "cudnn":
            raise PwcError("Trainer: the cuDNN baseline arm has no backward path")
        if len(weights) != model.output_level + 1:
            raise ValueError(f"need {model.output_level + 1} loss weights (one per pyramid flow), got {len(weights)}")
        self.model = model
        self.device = model.device
        self.lr, self.gamma = float(lr), float(gamma)
        self.loss_weights = [float(w) for w in weights]
        self.beta1, self.beta2, self.eps = float(beta1), float(beta2), float(eps)
        self.lr_boundaries = list(lr_boundaries)
        self.global_step = int(global_step)            # train.py:81
        self.pg = process_group
        dev = model.device
        n = model.flat.numel()
        self.grad_flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)      # <var>/Adam
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)      # <var>/Adam_1
        self.grads: Dict[str, torch.Tensor] = {}
        off = 0
        for name in model.var_names:
            t = model.params[name]
            self.grads[name] = self.grad_flat[off:off + t.numel()].view(t.shape)
            off += t.numel()
        self._var_off: Dict[str, tuple] = {}            # variable name -> (begin, end) in the flat buffers
        off = 0
        for name in model.var_names:
            cnt = model.params[name].numel()
            self._var_off[name] = (off, off + cnt)
            off += cnt
        # data parallel: the gradient all-reduce is issued in buckets as soon as a range of the flat gradient is final
        # (context + estimator level by level, then pyramid level by level, deepest first) on a side stream, so only the
        # last small bucket (pyramid level 1: 5 k floats) is exposed; skip_allreduce is a measurement switch (bench.py)
        self.skip_allreduce = False
        self._ar_stream = None
        self._ar_works: list = []
        self._lr_t = torch.zeros(1, dtype=torch.float32, device=dev)
        self._scalars = torch.zeros(3, dtype=torch.float32, device=dev)   # multiscale loss, l2 term, epe
        self._gbufs: Dict[tuple, _Grads] = {}
        self._parts: Dict[str, list] = {}
        self.tc_dgrad = model.precision == "3xf16" if tc_dgrad is None else bool(tc_dgrad)
        self.tc_wgrad = model.precision == "3xf16" if tc_wgrad is None else bool(tc_wgrad)
        self._tscratch = None
        self._dil = None
        self._dgrad_jobs, self._dgrad_jobs_n, self._dgrad_batched = None, -1, set()
        import os
        self.tc_dgrad_s2 = os.environ.get("PWC_DGRAD_S2_TC", "1") != "0"
        self.tc_wgrad_small = os.environ.get("PWC_WGRAD_TC_SMALL", "0") == "1"
        # weight gradients on a side stream: a layer's wgrad (transpose/split passes + GEMM) only needs x and dy, the chain
        # that the step waits for is dgrad -> dgrad; the bandwidth-bound transposes run under the tensor-bound dgrads
        self.wgrad_stream = os.environ.get("PWC_WGRAD_STREAM", "1") == "1"     # 14.6 -> 13.7 ms per step at B = 8, 384x1024
        self._ws = None

    # ------------------------------------------------------------------ gradient workspace
    def _grad_buffers(self, p) -> _Grads:
        key = (p.B, p.H, p.W)
        g = self._gbufs.get(key)
        if g is not None:
            return g
        # (activation, cleared every step?)  A buffer whose FIRST writer in `backward` accumulates (several consumers: pyramid
        # outputs, concat buffers, flows, the estimator's last buffer / dense buffer) is cleared; a buffer that one dgrad
        # overwrites completely (pyramid conv 0/1 outputs, estimator conv 0..3 outputs, context outputs, warped features) is
        # not: ~1.3 of the 1.5 GB at 8 x 384 x 1024.  The cleared ones come first in `flat`, so one fill covers them.
        acts: List[tuple] = []
        for lev in p.pyr:
            acts += [(a, j == len(lev) - 1) for j, a in enumerate(lev)]
        for l in range(len(p.S)):
            tmp = list(p.tmp[l] or [])
            acts += [(p.S[l], True)] + [(a, j == len(tmp) - 1) for j, a in enumerate(tmp)] + [(p.flows[l], True)] \
                + ([(p.f1w[l], True)] if p.f1w[l] is not None else [])
        acts += [(a, False) for a in p.ctx]
        def al(n):   # every view starts on a 256-byte boundary (vector loads / TMA need 16)
            return (n + 63) // 64 * 64
        total = sum(al(a.numel()) for a, _ in acts)
        g = _Grads()
        g.flat = torch.zeros(total, dtype=torch.float32, device=self.model.device)
        g.n_clear = sum(al(a.numel()) for a, z in acts if z)
        off_z, off_o = 0, g.n_clear
        views = []
        for a, z in acts:
            off = off_z if z else off_o
            views.append(g.flat[off:off + a.numel()].view(a.shape))
            if z:
                off_z += al(a.numel())
            else:
                off_o += al(a.numel())
        it = iter(views)
        g.pyr = [[next(it) for _ in lev] for lev in p.pyr]
        g.S, g.tmp, g.flows, g.f1w = [], [], [], []
        for l in range(len(p.S)):
            g.S.append(next(it))
            g.tmp.append([next(it) for _ in (p.tmp[l] or [])])
            g.flows.append(next(it))
            g.f1w.append(next(it) if p.f1w[l] is not None else None)
        g.ctx = [next(it) for _ in p.ctx]
        self._gbufs[key] = g
        return g

    # ------------------------------------------------------------------ one layer
    def _conv_bwd(self, scope, x, dy, gx, stride=1, dilation=1, mask=None, accumulate=False):
        """wgrad + bias grad into the flat gradient, then dgrad into gx (skipped when gx is None)."""
        m = self.model
        cin, cout = x.shape[3], dy.shape[3]
        if self.wgrad_stream:
            # everything enqueued so far (dy is final, grad_flat is cleared) happens before the side stream's work
            main = torch.cuda.current_stream(self.device)
            if self._ws is None:
                self._ws = torch.cuda.Stream(device=self.device)
            self._ws.wait_stream(main)
            with torch.cuda.stream(self._ws):
                self._wgrad(scope, x, dy, stride, dilation)
        else:
            self._wgrad(scope, x, dy, stride, dilation)
        self._dgrad(scope, x, dy, gx, stride, dilation, mask, accumulate)

    def _join_wgrad(self) -> None:
        """The current stream waits for the weight gradients issued on the side stream so far."""
        if self.wgrad_stream and self._ws is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._ws)

    def _wgrad(self, scope, x, dy, stride, dilation) -> None:
        m = self.model
        cin, cout = x.shape[3], dy.shape[3]
        # (layers with Cin, Cout <= 32 stay on the persistent CUDA-core kernel: measured equal, one launch instead of three)
        if self.tc_wgrad and cout % 16 == 0 and cin % 4 == 0 and (self.tc_wgrad_small or not (cin <= 32 and cout <= 32)) and stride in (1, 2) \
                and x.stride(2) % 4 == 0 and dy.stride(2) % 4 == 0 and x.data_ptr() % 16 == 0 and dy.data_ptr() % 16 == 0:
            # tensor-core wgrad: transpose + split both operands into channel-major fp16 planes (the dy pass also
            # reduces the bias gradient), then one GEMM launch over (tap pairs, channel tiles, row ranges)
            B, H, W = x.shape[0], x.shape[1], x.shape[2]
            need = max(ops_bwd.tsplit_bytes(B, H, dy.shape[2], cin, 3), ops_bwd.tsplit_bytes(B, dy.shape[1], dy.shape[2], cout)) // 2
            if self._tscratch is None or self._tscratch[0].numel() < need:
                self._tscratch = [torch.empty(need, dtype=torch.float16, device=x.device) for _ in range(2)]
            xT = ops_bwd.tsplit(x, out=self._tscratch[0], conv_input=True, stride=stride, dilation=dilation)
            dyT = ops_bwd.tsplit(dy, out=self._tscratch[1], db=self.grads[scope + "/bias"])
            ops_bwd.conv3x3_wgrad_tc(xT, dyT, self.grads[scope + "/kernel"], (B, H, W, cin), cout, stride=stride,
                                     dilation=dilation, cin_map=m._cin_perm.get(scope))
        else:
            ops_bwd.conv3x3_wgrad(x, dy, self.grads[scope + "/kernel"], self.grads[scope + "/bias"], stride=stride,
                                  dilation=dilation, cin_map=m._cin_perm.get(scope))

    def _dgrad(self, scope, x, dy, gx, stride, dilation, mask, accumulate) -> None:
        m = self.model
        cin, cout = x.shape[3], dy.shape[3]
        if gx is None:
            return
        k = m._k[scope]
        if self.tc_dgrad and self.tc_dgrad_s2 and stride == 2 and dilation == 1 and cout >= 16 and cout % 4 == 0 \
                and dy.stride(2) % 4 == 0 and dy.data_ptr() % 16 == 0:
            # stride-2 dgrad == stride-1 dgrad of the zero-inserted dy (odd/even positions chosen by the SAME padding):
            # one streaming pass + the tcgen05 kernel instead of the CUDA-core parity-class kernel (conv2a: 681 -> ~200 us)
            B, H, W = gx.shape[0], gx.shape[1], gx.shape[2]
            OH, OW = dy.shape[1], dy.shape[2]
            pad_t, pad_l = max((OH - 1) * 2 + 3 - H, 0) // 2, max((OW - 1) * 2 + 3 - W, 0) // 2
            n = B * H * W * cout
            if self._dil is None or self._dil.numel() < n:
                self._dil = torch.empty(n, dtype=torch.float32, device=dy.device)
            dy = ops_bwd.dilate2(dy, self._dil[:n].view(B, H, W, cout), 1 - pad_t, 1 - pad_l)
            stride = 1
        if self.tc_dgrad and stride == 1 and k.shape[3] >= 16 and dy.stride(2) % 4 == 0 and dy.data_ptr() % 16 == 0:
            # stride-1 dgrad == SAME conv of dy with the rotated kernel: runs on the tcgen05 forward kernel
            # (3 x fp16 split, fp32-class).  dx channel ranges wider than 256 / not a multiple of 16 are split / padded.
            from . import ops_tc
            batched = scope in self._dgrad_batched      # packed by the one launch at the start of backward()
            for c0, cnt, pad, rot, packed in self._dgrad_parts(scope):
                if not batched:
                    ops_bwd.rot_weights(k, out=rot, ci_begin=c0, ci_count=cnt, ci_pad=pad)
                    ops_tc.pack_weights_f16(rot, out=packed)
                ops_bwd.conv3x3_tc_f16_dgrad(dy, packed, gx[..., c0:c0 + cnt], pad, dilation=dilation,
                                             mask=None if mask is None else mask[..., c0:c0 + cnt], mask_alpha=0.1,
                                             accumulate=accumulate)
        else:
            ops_bwd.conv3x3_dgrad(dy, k, gx, stride=stride, dilation=dilation, mask=mask, mask_alpha=0.1,
                                  accumulate=accumulate)

    def _dgrad_parts(self, scope):
        parts = self._parts.get(scope)
        if parts is None:
            from ._abi import lib
            k = self.model._k[scope]
            cdx, cdy = k.shape[2], k.shape[3]
            n_parts = -(-cdx // 128)          # parts of <= 128 channels run on the halo-resident kernel
            size = (-(-cdx // n_parts) + 15) // 16 * 16
            parts = []
            for c0 in range(0, cdx, size):
                cnt = min(size, cdx - c0)
                pad = (cnt + 15) // 16 * 16
                rot = torch.empty((3, 3, cdy, pad), dtype=torch.float32, device=k.device)
                packed = torch.empty(lib().pwc_conv3x3_packed_bytes_f16(cdy, pad) // 2, dtype=torch.float16, device=k.device)
                parts.append((c0, cnt, pad, rot, packed))
            self._parts[scope] = parts
        return parts

    # ------------------------------------------------------------------ backward
    @on_device
    def backward(self, p, flows_gt) -> None:
        """Gradient of multiscale_loss(flows_gt, flows_pyramid) w.r.t. every variable -> self.grad_flat
        (the regulariser's gradient gamma*var is added inside the Adam kernel)."""
        m = self.model
        n, B, nd, L = m.name, p.B, m._nd, m.output_level
        nest = len(ESTIMATOR_FILTERS)
        nf = ESTIMATOR_FILTERS[-1]
        g = self._grad_buffers(p)
        g.flat[:g.n_clear].zero_()          # only the buffers whose first writer accumulates (see _grad_buffers)
        self.grad_flat.zero_()
        # rotated + packed dgrad kernels of every layer seen by an earlier backward pass: one launch (the first pass
        # packs layer by layer and records the parts)
        if self._dgrad_jobs is None or self._dgrad_jobs_n != len(self._parts):
            from . import ops_tc
            self._dgrad_jobs = ops_tc.PackJobs(m.device)
            for scope, parts in self._parts.items():
                for c0, cnt, pad, rot, packed in parts:
                    self._dgrad_jobs.add_dgrad(m._k[scope], packed, c0, cnt, pad)
            self._dgrad_jobs_n = len(self._parts)
            self._dgrad_batched = set(self._parts)
        self._dgrad_jobs.run()
        for l in range(L + 1):
            ops_bwd.lploss_level_bwd(flows_gt, p.flows[l], self.loss_weights[l], g.flows[l], gt_div=20.0, ord=2)

        pre_total = sum(ESTIMATOR_FILTERS) if m.use_dc else 0
        for l in range(L, -1, -1):
            lv = m._lv[l]
            lev = m.num_levels - 1 - l
            S, gS = p.S[l], g.S[l]
            head = f"{n}/optflow_{l}/conv2d_{nest}"
            if m.use_dc:
                # ---- dense connections (modules.py:269-270): one buffer per level, [conv4 | conv3 | conv2 | conv1 | conv0 | X];
                # conv i reads channels [start_i, end) and writes [start_i - f_i, start_i).  Every region collects the
                # gradients of ALL its consumers (later convs, head, x2 up-sampling, context) before its own conv runs.
                end = pre_total + lv["cin_int"]
                feats, gfeats = S[..., 0:end], gS[..., 0:end]
                X, gX = S[..., pre_total:end], gS[..., pre_total:end]
                flow_slot_g = gX[..., lv["off_flow"]:lv["off_flow"] + 2] if l else None
                if l == L:
                    nctx = len(CONTEXT_FILTERS)
                    for i in range(nctx - 1, 0, -1):
                        scope = f"{n}/context/conv2d_{i}"
                        dy = g.flows[l] if i == nctx - 1 else g.ctx[i]
                        self._conv_bwd(scope, p.ctx[i - 1], dy, g.ctx[i - 1], dilation=CONTEXT_DILATIONS[i], mask=p.ctx[i - 1])
                    ops_bwd.add_(gS[..., end:end + 2], g.flows[l])          # residual `flows + x` (modules.py:326)
                    self._conv_bwd(f"{n}/context/conv2d", S[..., 0:end + 4], g.ctx[0], gS[..., 0:end + 4],
                                   dilation=CONTEXT_DILATIONS[0], accumulate=True)
                    dy_head = gS[..., end:end + 2]
                else:
                    dy_head = g.flows[l]
                self._conv_bwd(head, feats, dy_head, gfeats, accumulate=True)
                if l:
                    ops_bwd.add_(flow_slot_g, dy_head)
                start = 0
                for i in range(nest - 1, -1, -1):
                    f = ESTIMATOR_FILTERS[i]
                    scope = f"{n}/optflow_{l}/conv2d" + (f"_{i}" if i else "")
                    out_g, out_y = gS[..., start:start + f], S[..., start:start + f]
                    ops_bwd.leaky_bwd(out_g, out_y, 0.1)
                    self._conv_bwd(scope, S[..., start + f:end], out_g, gS[..., start + f:end], accumulate=True)
                    start += f
                gS_in, S_in = gX, X
                up_feat_g = lambda lprev: g.S[lprev][..., 0:pre_total + m._lv[lprev]["cin_int"]]
            else:
                feats, gfeats = p.tmp[l][-1][..., 0:nf], g.tmp[l][-1][..., 0:nf]
                flow_slot_g = gS[..., lv["off_flow"]:lv["off_flow"] + 2] if l else None
                if l == L:
                    # ---- context network (modules.py:304-326): flows_out = flow_slot + conv6(...conv0([features, flows]))
                    Cbuf, gCbuf = p.tmp[l][-1], g.tmp[l][-1]            # [features nf | flows 2 | pad 2]
                    nctx = len(CONTEXT_FILTERS)
                    for i in range(nctx - 1, 0, -1):
                        scope = f"{n}/context/conv2d_{i}"
                        dy = g.flows[l] if i == nctx - 1 else g.ctx[i]
                        self._conv_bwd(scope, p.ctx[i - 1], dy, g.ctx[i - 1], dilation=CONTEXT_DILATIONS[i], mask=p.ctx[i - 1])
                    ops_bwd.add_(gCbuf[..., nf:nf + 2], g.flows[l])     # residual `flows + x` (modules.py:326)
                    self._conv_bwd(f"{n}/context/conv2d", Cbuf[..., 0:nf + 4], g.ctx[0], gCbuf[..., 0:nf + 4],
                                   dilation=CONTEXT_DILATIONS[0], accumulate=True)
                    dy_head = gCbuf[..., nf:nf + 2]
                else:
                    dy_head = g.flows[l]
                # ---- flow head (no activation) + residual flows_up (modules.py:274-277)
                self._conv_bwd(head, feats, dy_head, gfeats, accumulate=True)
                if l:
                    ops_bwd.add_(flow_slot_g, dy_head)
                ops_bwd.leaky_bwd(gfeats, feats, 0.1)
                # ---- estimator convs (modules.py:266-270)
                for i in range(nest - 1, 0, -1):
                    scope = f"{n}/optflow_{l}/conv2d_{i}"
                    fi = ESTIMATOR_FILTERS[i]
                    self._conv_bwd(scope, p.tmp[l][i - 1], g.tmp[l][i][..., 0:fi], g.tmp[l][i - 1], mask=p.tmp[l][i - 1])
                self._conv_bwd(f"{n}/optflow_{l}/conv2d", S, g.tmp[l][0], gS, accumulate=True)
                gS_in, S_in = gS, S
                up_feat_g = lambda lprev: g.tmp[lprev][-1][..., 0:nf]
            # ---- concat slots: cost volume (+ f0 copy), warp, x2 up-sampling (model.py:106-112, modules.py:262-285)
            F, gF = p.pyr[lev][2], g.pyr[lev][2]
            f0, f1 = F[:B], F[B:]
            g_cv, cv = gS_in[..., 0:nd], S_in[..., 0:nd]
            g_f0 = gS_in[..., lv["off_f0"]:lv["off_f0"] + lv["C"]]
            if l == 0:
                ops_bwd.cost_volume_bwd(g_cv, cv, f0, f1, gF[:B], gF[B:], g_f0slot=g_f0, accumulate_f1=True,
                                        search_range=m.s_range)
            else:
                ops_bwd.cost_volume_bwd(g_cv, cv, f0, p.f1w[l], gF[:B], g.f1w[l], g_f0slot=g_f0, accumulate_f1=False,
                                        search_range=m.s_range)
                flow_up = S_in[..., lv["off_flow"]:lv["off_flow"] + 2]
                ops_bwd.warp_bwd(f1, flow_up, g.f1w[l], gF[B:], dflow=flow_slot_g, flow_scale=m.scales[l],
                                 warp_type=m.warp_type)
                ops_bwd.resize_bilinear_bwd(flow_slot_g, g.flows[l - 1])
                up = lv["up"]
                ops_bwd.resize_bilinear_bwd(gS_in[..., lv["off_feat"]:lv["off_feat"] + up], up_feat_g(l - 1))
            # this level's estimator (and, at the output level, the context network) gradients are final
            last = f"{n}/context/conv2d_{len(CONTEXT_FILTERS) - 1}/bias" if l == L else f"{n}/optflow_{l}/conv2d_{nest}/bias"
            self._allreduce_range(f"{n}/optflow_{l}/conv2d/kernel", last)

        # ---- feature pyramid, both images as one batch of 2B (modules.py:49-71)
        for lev in range(m.num_levels - 1, -1, -1):
            def scope(j, lev=lev):
                idx = 3 * lev + j
                return f"{n}/fp_extractor/conv2d" + (f"_{idx}" if idx else "")
            ops_bwd.leaky_bwd(g.pyr[lev][2], p.pyr[lev][2], 0.1)
            self._conv_bwd(scope(2), p.pyr[lev][1], g.pyr[lev][2], g.pyr[lev][1], mask=p.pyr[lev][1])
            self._conv_bwd(scope(1), p.pyr[lev][0], g.pyr[lev][1], g.pyr[lev][0], mask=p.pyr[lev][0])
            if lev:
                self._conv_bwd(scope(0), p.pyr[lev - 1][2], g.pyr[lev][0], g.pyr[lev - 1][2], stride=2, accumulate=True)
            else:
                self._conv_bwd(scope(0), p.im, g.pyr[lev][0], None, stride=2)
            self._allreduce_range(scope(0) + "/kernel", scope(2) + "/bias")
        self._join_wgrad()

    # ------------------------------------------------------------------ gradient all-reduce, bucketed
    def _world(self) -> int:
        import torch.distributed as dist
        if self.skip_allreduce or not (dist.is_available() and dist.is_initialized()):
            return 1
        return dist.get_world_size(self.pg)

    def _allreduce_range(self, first_var: str, last_var: str) -> None:
        """SUM all-reduce of grad_flat[begin(first_var) : end(last_var)] on the side stream, after everything enqueued so far."""
        if self._world() == 1:
            return
        import torch.distributed as dist
        lo, hi = self._var_off[first_var][0], self._var_off[last_var][1]
        cur = torch.cuda.current_stream(self.device)
        if self._ar_stream is None:
            self._ar_stream = torch.cuda.Stream(device=self.device)
        ev = torch.cuda.Event()
        ev.record(cur)
        ev_w = None
        if self.wgrad_stream and self._ws is not None:      # the range's weight gradients were issued on the side stream:
            ev_w = torch.cuda.Event()                       # the collective waits for them, the dgrad chain does not
            ev_w.record(self._ws)
        with torch.cuda.stream(self._ar_stream):
            self._ar_stream.wait_event(ev)
            if ev_w is not None:
                self._ar_stream.wait_event(ev_w)
            self._ar_works.append(dist.all_reduce(self.grad_flat[lo:hi], op=dist.ReduceOp.SUM, group=self.pg, async_op=True))

    def _allreduce_finish(self) -> int:
        """Wait (on the current stream) for the buckets; all-reduce the three logged scalars.  Returns the world size."""
        world = self._world()
        if world == 1:
            return 1
        import torch.distributed as dist
        for w in self._ar_works:
            w.wait()
        self._ar_works.clear()
        dist.all_reduce(self._scalars, op=dist.ReduceOp.SUM, group=self.pg)
        self._scalars /= world
        return world

    # ------------------------------------------------------------------ losses
    def _losses(self, p, flows_gt) -> None:
        s = self._scalars
        s.zero_()
        for w, fs in zip(self.loss_weights, p.flows):
            ops.lploss_level(flows_gt, fs, w, s[0:1], gt_div=20.0, ord=2)     # losses.py:15-31
        ops_bwd.sumsq(self.model.flat, s[1:2], 0.5)                           # train.py:74
        ops.epe(flows_gt, p.flows_final, s[2:3])                              # train.py:77

    # ------------------------------------------------------------------ step
    def lr_t(self, t: int) -> float:
        """Learning rate TF's Adam applies at its t-th update: lr(global_step) * sqrt(1-b2^t)/(1-b1^t)."""
        lr = piecewise_lr(self.global_step, self.lr, self.lr_boundaries)
        return lr * math.sqrt(1.0 - self.beta2 ** t) / (1.0 - self.beta1 ** t)

    @on_device
    def forward_backward(self, images_0, images_1, flows_gt):
        m = self.model
        i0, i1 = m._as_input(images_0, "images_0"), m._as_input(images_1, "images_1")
        if i0.shape != i1.shape:
            raise ValueError("images_0 and images_1 differ in shape")
        B, H, W, C = i0.shape
        m._check_shape(B, H, W, C)
        gt = flows_gt if isinstance(flows_gt, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(flows_gt))
        if gt.dtype != torch.float32 or tuple(gt.shape) != (B, H, W, 2):
            raise ValueError(f"flows_gt must be float32 (B,H,W,2) = {(B, H, W, 2)}, got {tuple(gt.shape)} {gt.dtype}")
        gt = gt.to(m.device, non_blocking=True).contiguous()
        if i0.dtype != i1.dtype:
            raise ValueError("images_0 and images_1 differ in dtype")
        p = m.plan(B, H, W)
        m._stage(p, i0, i1)              # float32 RGB/255 or uint8 bytes (train.py:122's /255.0 then happens on the device)
        m._launch(p)
        self._losses(p, gt)
        self.backward(p, gt)
        return p

    @on_device
    def step(self, images_0, images_1, flows_gt):
        """One optimisation step (train.py:143-147).  Returns (loss, multiscale_loss, epe) as 0-dim CUDA tensors;
        with a process group the three scalars are averaged over ranks like the gradient."""
        self.forward_backward(images_0, images_1, flows_gt)
        world = self._allreduce_finish()                 # the one data-path collective: buckets issued during backward()
        t = self.global_step + 1
        self._lr_t.fill_(self.lr_t(t))
        ops_bwd.adam_step(self.model.flat, self.grad_flat, self.m, self.v, self._lr_t, self.beta1, self.beta2, self.eps,
                          gamma=self.gamma, grad_scale=1.0 / world)
        self.global_step += 1                                                         # train.py:91-92
        self.model.refresh_derived()
        s = self._scalars
        return s[0] + self.gamma * s[1], s[0].clone(), s[2].clone()

    @on_device
    def evaluate(self, images_0, images_1, flows_gt):
        """Forward + losses without an update: what `sess.run(self.merged)` computes for the summaries and the
        validation pass (train.py:128-141).  Returns (loss, multiscale_loss, epe) like step()."""
        m = self.model
        i0, i1 = m._as_input(images_0, "images_0"), m._as_input(images_1, "images_1")
        B, H, W, C = i0.shape
        m._check_shape(B, H, W, C)
        gt = flows_gt if isinstance(flows_gt, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(flows_gt))
        if gt.dtype != torch.float32 or tuple(gt.shape) != (B, H, W, 2):
            raise ValueError(f"flows_gt must be float32 (B,H,W,2) = {(B, H, W, 2)}, got {tuple(gt.shape)} {gt.dtype}")
        gt = gt.to(m.device, non_blocking=True).contiguous()
        p = m.plan(B, H, W)
        m._stage(p, i0, i1)
        m._launch(p)
        self._losses(p, gt)
        s = self._scalars
        return s[0] + self.gamma * s[1], s[0].clone(), s[2].clone()

    # ------------------------------------------------------------------ checkpoints (train.py:95-99,166)
    def state_dict(self) -> Dict[str, np.ndarray]:
        """Everything tf.train.Saver stores for the reference's training graph, under the same names: the 110
        variables, their Adam slots `<var>/Adam` (m) and `<var>/Adam_1` (v), `Variable` (global_step, int32),
        `beta1_power`, `beta2_power`."""
        m = self.model
        sd = m.state_dict()
        off = 0
        mh, vh = self.m.cpu().numpy(), self.v.cpu().numpy()
        for name in m.var_names:
            t = m.params[name]
            n = t.numel()
            sd[name + "/Adam"] = mh[off:off + n].reshape(tuple(t.shape)).copy()
            sd[name + "/Adam_1"] = vh[off:off + n].reshape(tuple(t.shape)).copy()
            off += n
        sd["Variable"] = np.array(self.global_step, np.int32)
        sd["beta1_power"] = np.array(self.beta1 ** (self.global_step + 1), np.float32)   # TF stores beta^(t+1) after t updates
        sd["beta2_power"] = np.array(self.beta2 ** (self.global_step + 1), np.float32)
        return sd

    @on_device
    def load_state_dict(self, sd) -> None:
        """Inverse of state_dict(); `sd` may also be the prefix of a checkpoint bundle -- written by save() below or by
        the reference's tf.train.Saver (e.g. 'model_250.ckpt') -- or the path of a legacy `.npz` written by round-1
        builds: weights, Adam moments and global_step are resumed.  A bundle that holds only the model variables
        (train.py's current `Saver(model.vars)`, no 'Variable' / Adam slots) resumes with global_step 0 and zero
        moments, which is what the reference does on `--resume`.  The Adam bias correction is derived from global_step
        (t = global_step + 1); stored beta powers are not read."""
        if isinstance(sd, str):
            if sd.endswith(".npz") and os.path.exists(sd):
                sd = dict(np.load(sd))
            else:
                from .checkpoint import load_checkpoint, list_variables, read_scalar
                prefix = sd
                sd = load_checkpoint(prefix, self.model.name, include_slots=True)
                names = list_variables(prefix)
                for gs in ("Variable", "global_step"):
                    if gs in names:
                        sd["Variable"] = read_scalar(prefix, gs)
                        break
        m = self.model
        m.load_weights({k: sd[k] for k in m.var_names})
        off = 0
        for name in m.var_names:
            n = m.params[name].numel()
            for buf, sfx in ((self.m, "/Adam"), (self.v, "/Adam_1")):
                if name + sfx in sd:
                    buf[off:off + n].copy_(torch.from_numpy(np.ascontiguousarray(sd[name + sfx], np.float32)).reshape(-1))
                else:
                    buf[off:off + n].zero_()
            off += n
        self.global_step = int(np.asarray(sd["Variable"]).reshape(-1)[0]) if "Variable" in sd else 0

    def save(self, path: str) -> None:
        """Per-epoch checkpoint (train.py:164-166, `saver.save(sess, 'model_{e+1}.ckpt')`): a TensorFlow checkpoint
        bundle `<path>.index` + `<path>.data-00000-of-00001` under the reference's variable names, byte-compatible with
        tf.train.Saver (pwcnet_b200/checkpoint.py:save_checkpoint), so the reference's test.py / train.py --resume and
        this package's load_weights / load_state_dict / `infer.py --resume` all read it."""
        from .checkpoint import save_checkpoint
        save_checkpoint(path, self.state_dict())

    @on_device
    def launches_per_step(self) -> int:
        """Kernel launches of ours in one training step, counted: C-ABI calls issued by one step (losses, backward,
        optimizer, derived-weight refresh) plus the forward's launches (replayed from its CUDA graph)."""
        from . import _abi
        if not self._gbufs:
            raise RuntimeError("launches_per_step() needs one step to have run")
        p = next(pl for pl in self.model._plans.values() if not pl.u8)
        gt = torch.zeros((p.B, p.H, p.W, 2), dtype=torch.float32, device=self.model.device)
        n0 = _abi.LAUNCHES
        self._losses(p, gt)
        self.backward(p, gt)
        n1 = _abi.LAUNCHES
        upd = 1 + len(self.model._cin_perm) + len(self.model._packed)
        return self.model.launches_per_forward() + (n1 - n0) + upd
