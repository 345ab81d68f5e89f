"""Host<->device pipelining for throughput inference through the public model API.

`PWCDCNet.__call__` on host arrays is synchronous in effect (H2D, forward, then the caller's D2H).
`InferenceStream` keeps `depth` requests in flight: the H2D copy of request i+1 (copy engine, its own
stream) and the D2H copy of request i-1 overlap the forward of request i, so end-to-end throughput is
max(compute, PCIe) instead of their sum.  Inputs come from (pinned) host memory every request and the
results land in pinned host buffers: same data movement as the synchronous call.

    stream = InferenceStream(model, depth=2)
    tickets = [stream.submit(im0, im1) for ...]      # host float32 (B,H,W,3) arrays / tensors
    flows_final, flows_pyramid = stream.collect(ticket)   # torch CPU tensors (pinned), valid until the slot is reused
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch


class InferenceStream:
    """`outputs` selects what travels back to the host: "final" (flows_final, full resolution) and/or "pyramid" (the five
    per-scale flows -- all the reference's test.py fetches, test.py:51).  Images may be float32 RGB/255 or uint8 RGB bytes
    (then the reference's host-side `/255.0` runs on the device, bit-identically, and a quarter of the bytes cross PCIe)."""

    def __init__(self, model, depth: int = 2, outputs=("final", "pyramid")):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        outputs = tuple(outputs)
        if not outputs or any(o not in ("final", "pyramid") for o in outputs):
            raise ValueError("outputs must be a non-empty subset of ('final', 'pyramid')")
        self.model = model
        self.depth = depth
        self.outputs = outputs
        self.dev = model.device
        self.s_in = torch.cuda.Stream(device=self.dev)
        self.s_out = torch.cuda.Stream(device=self.dev)
        self._slots = None
        self._shape = None
        self._next = 0
        self._pending = {}

    def _alloc(self, B, H, W, dtype):
        plan = self.model.plan(B, H, W, u8=dtype == torch.uint8 and self.model.precision != "cudnn")
        outs = ([plan.flows_final] if "final" in self.outputs else []) + (list(plan.flows) if "pyramid" in self.outputs else [])
        self._slots = []
        for _ in range(self.depth):
            self._slots.append(dict(
                in_dev=torch.empty((2 * B, H, W, 3), dtype=dtype, device=self.dev),
                out_dev=[torch.empty_like(t) for t in outs],
                out_host=[torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in outs],
                nf_dev=torch.zeros(1, dtype=torch.int32, device=self.dev), nf_host=torch.zeros(1, dtype=torch.int32).pin_memory(),
                ev_in=torch.cuda.Event(), ev_taken=torch.cuda.Event(), ev_fwd=torch.cuda.Event(),
                ev_out=torch.cuda.Event(), used=False, ticket=None))
        self._shape = (B, H, W, dtype)

    @staticmethod
    def _host(a, name):
        if isinstance(a, np.ndarray):
            if a.dtype not in (np.float32, np.uint8):
                raise TypeError(f"{name}: dtype must be float32 or uint8")
            a = torch.from_numpy(np.ascontiguousarray(a))
        if not isinstance(a, torch.Tensor) or a.is_cuda or a.dtype not in (torch.float32, torch.uint8) or a.dim() != 4:
            raise TypeError(f"{name}: expected a float32 or uint8 host array/tensor of shape (B,H,W,3)")
        return a

    def submit(self, images_0, images_1) -> int:
        i0, i1 = self._host(images_0, "images_0"), self._host(images_1, "images_1")
        if i0.shape != i1.shape or i0.dtype != i1.dtype:
            raise ValueError("images_0 and images_1 differ in shape or dtype")
        B, H, W, C = i0.shape
        with torch.cuda.device(self.dev):
            if self._shape != (B, H, W, i0.dtype):
                self.model._check_shape(B, H, W, C)          # before any allocation
                if self._pending:
                    raise RuntimeError("submit: input shape/dtype changed while tickets are pending; collect() them first")
                self._alloc(B, H, W, i0.dtype)
            ticket = self._next
            self._next += 1
            sl = self._slots[ticket % self.depth]
            if sl["used"] and sl["ticket"] in self._pending:
                raise RuntimeError(f"slot still holds uncollected ticket {sl['ticket']}: collect() it first (depth={self.depth})")
            cur = torch.cuda.current_stream(self.dev)
            # ---- H2D on the input stream (waits until the previous user of this slot was consumed)
            with torch.cuda.stream(self.s_in):
                if sl["used"]:
                    self.s_in.wait_event(sl["ev_taken"])
                sl["in_dev"][:B].copy_(i0, non_blocking=True)
                sl["in_dev"][B:].copy_(i1, non_blocking=True)
                sl["ev_in"].record(self.s_in)
            # ---- forward on the caller's stream
            cur.wait_event(sl["ev_in"])
            flows_final, pyr = self.model._run_device(sl["in_dev"], B, H, W)
            sl["ev_taken"].record(cur)
            if sl["used"]:
                cur.wait_event(sl["ev_out"])            # previous D2H out of this slot's staging has finished
            srcs = ([flows_final] if "final" in self.outputs else []) + (list(pyr) if "pyramid" in self.outputs else [])
            for d, s in zip(sl["out_dev"], srcs):
                d.copy_(s, non_blocking=True)
            sl["nf_dev"].copy_(self.model.plan(B, H, W, u8=i0.dtype == torch.uint8 and self.model.precision != "cudnn").nonfinite, non_blocking=True)      # range guard of this request
            sl["ev_fwd"].record(cur)
            # ---- D2H on the output stream
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(sl["ev_fwd"])
                for h, d in zip(sl["out_host"], sl["out_dev"]):
                    h.copy_(d, non_blocking=True)
                sl["nf_host"].copy_(sl["nf_dev"], non_blocking=True)
                sl["ev_out"].record(self.s_out)
            sl["used"], sl["ticket"] = True, ticket
            self._pending[ticket] = sl
        return ticket

    def collect(self, ticket: int):
        """-> (flows_final, flows_pyramid) as pinned host tensors (valid until the slot is reused); an entry that was not
        requested through `outputs` is None."""
        sl = self._pending.pop(ticket)
        sl["ev_out"].synchronize()
        if int(sl["nf_host"][0]) != 0:
            from ._abi import PwcError
            raise PwcError(f"ticket {ticket}: {int(sl['nf_host'][0])} non-finite flow values: activations or weights left the fp16 "
                           f"range of precision='{self.model.precision}' (see PWCDCNet.check_finite)")
        oh = sl["out_host"]
        final = oh[0] if "final" in self.outputs else None
        pyr = (oh[1:] if "final" in self.outputs else oh) if "pyramid" in self.outputs else None
        return final, pyr

    def drain(self):
        for t in list(self._pending):
            self.collect(t)


class TrainStream:
    """Input side of the training loop (the reference feeds its step from a torch DataLoader with worker processes and
    pinned batches, train.py:36-41): `submit()` starts the H2D copy of a (pinned) host batch into one of `depth` device staging sets on
    a copy stream, `step()` runs `Trainer.step` on the oldest staged batch.  With depth >= 2 the copy of batch i+1
    overlaps the step of batch i; the bytes moved per step are those of the synchronous `Trainer.step(host...)`.

        ts = TrainStream(trainer)
        ts.submit(im0, im1, gt)
        for batch in batches:            # steady state: one submit, one step
            ts.submit(*batch)
            loss, multiscale_loss, epe = ts.step()
    """

    def __init__(self, trainer, depth: int = 2):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.trainer = trainer
        self.depth = depth
        self.dev = trainer.model.device
        self.s_in = torch.cuda.Stream(device=self.dev)
        self._slots = None
        self._shape = None
        self._head = 0          # next slot to fill
        self._tail = 0          # next slot to train on

    def _alloc(self, B, H, W, dtype=torch.float32):
        self._slots = [dict(im0=torch.empty((B, H, W, 3), dtype=dtype, device=self.dev),
                            im1=torch.empty((B, H, W, 3), dtype=dtype, device=self.dev),
                            gt=torch.empty((B, H, W, 2), dtype=torch.float32, device=self.dev),
                            ev_in=torch.cuda.Event(), ev_done=torch.cuda.Event(), used=False)
                       for _ in range(self.depth)]
        self._shape = (B, H, W, dtype)
        self._head = self._tail = 0

    def pending(self) -> int:
        return self._head - self._tail

    def submit(self, images_0, images_1, flows_gt) -> None:
        i0, i1 = InferenceStream._host(images_0, "images_0"), InferenceStream._host(images_1, "images_1")
        gt = flows_gt if isinstance(flows_gt, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(flows_gt))
        B, H, W, _ = i0.shape
        if i1.shape != i0.shape or i1.dtype != i0.dtype or gt.is_cuda or gt.dtype != torch.float32 or tuple(gt.shape) != (B, H, W, 2):
            raise ValueError("submit: images_0/images_1 (B,H,W,3; float32 or uint8) and flows_gt (B,H,W,2) float32 must be host arrays of one batch")
        if self._shape != (B, H, W, i0.dtype):
            if self.pending():
                raise RuntimeError("submit: batch shape changed while staged batches are pending; step() them first")
            self._alloc(B, H, W, i0.dtype)
        if self.pending() >= self.depth:
            raise RuntimeError(f"submit: all {self.depth} staging sets hold batches that were not trained on yet")
        sl = self._slots[self._head % self.depth]
        with torch.cuda.stream(self.s_in):
            if sl["used"]:
                self.s_in.wait_event(sl["ev_done"])      # the step that read this staging set has finished
            sl["im0"].copy_(i0, non_blocking=True)
            sl["im1"].copy_(i1, non_blocking=True)
            sl["gt"].copy_(gt, non_blocking=True)
            sl["ev_in"].record(self.s_in)
        sl["used"] = True
        self._head += 1

    def discard(self) -> None:
        """Drop the oldest staged batch without training on it (its staging set becomes reusable once the copy has landed)."""
        if not self.pending():
            raise RuntimeError("discard: no staged batch")
        sl = self._slots[self._tail % self.depth]
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(sl["ev_in"])
        sl["ev_done"].record(cur)
        self._tail += 1

    def step(self):
        if not self.pending():
            raise RuntimeError("step: no staged batch; submit() one first")
        sl = self._slots[self._tail % self.depth]
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(sl["ev_in"])
        out = self.trainer.step(sl["im0"], sl["im1"], sl["gt"])
        sl["ev_done"].record(cur)
        self._tail += 1
        return out
