#!/usr/bin/env python
"""bench.py -- headline benchmark of the PWC-Net hot path on B200.

Metric (BASELINE.json): image-pairs/sec at 448x1024 (PWCDCNet(use_dc=False) inference, the runnable
"PWCNet" of the reference), per-rank batch 8 synthetic pairs (BASELINE config 2 shape), data-parallel
over N GPUs with no collective on the data path ("scaling": "weak"); plus the level-2 cost-volume
kernel's achieved HBM GB/s against the measured copy peak (`roofline`), and the reference's CPU path
(restated on torch-CPU, TensorFlow 1.8 is not installable) timed beside it (`cpu_baseline`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--precision fp32|3xtf32|tf32|cudnn]
    python bench.py --impl reference ...      # the reference arm: CPU oracle on the box's host cores

One JSON line on stdout (rank 0).  Nothing here reads /root/reference.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 448, 1024
METRIC = "image-pairs/sec at 448x1024"
UNIT = "pairs/s"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _tensor_peak():
    """Dense bf16/fp16 TFLOP/s from MEASURED_PEAKS.json: the burst figure (the conv kernel is timed alone, one launch
    between L2 flushes), else the nominal number."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        for k in ("bf16_tflops", "bf16_tflops_sustained"):
            if k in d:
                return float(d[k]), f"measured (MEASURED_PEAKS.json {k})"
        vals = [float(v) for k, v in d.items() if "bf16" in k.lower() and isinstance(v, (int, float))]
        if vals:
            return min(vals), "measured (MEASURED_PEAKS.json, lowest bf16 figure)"
    return 2250.0, "nominal dense bf16 (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_oracle_pairs_per_s(n_pairs: int, warm: int = 1):
    """The reference's CPU path restated on torch-CPU (oracle/pwc_oracle.py), all host threads,
    one 448x1024 pair per forward (the reference's own `--time` loop runs batch 1, test.py:48-53)."""
    import torch
    from oracle import pwc_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Wt = O.glorot_weights(2)
    im0, im1 = O.synthetic_pair(1, H, W, 0)
    with torch.no_grad():
        for _ in range(warm):
            O.pwcdcnet_forward(Wt, im0, im1)
        times = []
        for _ in range(n_pairs):
            t = time.perf_counter()
            O.pwcdcnet_forward(Wt, im0, im1)
            times.append(time.perf_counter() - t)
    return 1.0 / float(np.mean(times)), cores, times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = max(args.steps, 1)
    for _ in range(max(args.warmup - 1, 0)):
        pass
    pps, cores, times = cpu_oracle_pairs_per_s(n, warm=max(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": pps, "unit": UNIT, "n_gpus": args.gpus, "steps": n,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "PWCDCNet(use_dc=False) inference 448x1024 (BASELINE config 2 shape)",
                   "sample": "each step = 1 pair (bounded sample of the 8-pair batch)"},
        "cpu_baseline": {"value": pps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n} forward passes of one 448x1024 pair, restated reference on torch-CPU "
                                   "(not TensorFlow 1.8: not installable here)"},
        "e2e": {"value": pps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_oracle_train_pairs_per_s(n_steps: int, h: int, w: int, warm: int = 1):
    """One reference training step (train.py:65-92) on the CPU oracle: forward, multiscale loss + regulariser,
    autograd backward, TF-Adam over the 110 variables; one pair per step (bounded sample)."""
    import torch
    from oracle import pwc_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    W0 = O.glorot_weights(2)
    Wt = {k: torch.from_numpy(v.copy()).requires_grad_(True) for k, v in W0.items()}
    M = {k: np.zeros_like(v) for k, v in W0.items()}
    V = {k: np.zeros_like(v) for k, v in W0.items()}
    im0, im1 = O.synthetic_pair(1, h, w, 0)
    gt = np.random.default_rng(1).normal(0, 5, (1, h, w, 2)).astype(np.float32)
    times = []
    for t in range(1, warm + n_steps + 1):
        t0 = time.perf_counter()
        total, epe, _, _ = O.training_loss(Wt, im0, im1, gt)
        total.backward()
        with torch.no_grad():
            for k, v in Wt.items():
                nv, M[k], V[k] = O.adam_step_tf(v.numpy(), v.grad.numpy(), M[k], V[k], t, O.piecewise_lr(t - 1))
                v.copy_(torch.from_numpy(nv)); v.grad = None
        if t > warm:
            times.append(time.perf_counter() - t0)
    return 1.0 / float(np.mean(times)), cores, times


TRAIN_H, TRAIN_W = 384, 1024   # factor_crop(436x1024, 64) (test.py:13-17): the Sintel shape the reference can execute
TRAIN_METRIC = "training image-pairs/sec at 384x1024 (Sintel shape cropped to /64)"


def run_train_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    n = max(min(args.steps, 5), 1)
    pps, cores, times = cpu_oracle_train_pairs_per_s(n, TRAIN_H, TRAIN_W, warm=1)
    line = {"impl": "reference", "metric": TRAIN_METRIC, "value": pps, "unit": UNIT, "n_gpus": args.gpus, "steps": n,
            "warmup": 1, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "PWCDCNet training step 384x1024 (BASELINE config 5)", "sample": "each step = 1 pair"},
            "cpu_baseline": {"value": pps, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{n} training steps on one 384x1024 pair, restated reference + torch autograd on CPU"},
            "e2e": {"value": pps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_train(args):
    """BASELINE config 5: PWCNet training step, batch 8/GPU, synthetic Sintel-shape pairs + random GT flow,
    multiscale L2 loss + regulariser, TF-Adam, NCCL all-reduce of the flat gradient at N > 1."""
    import torch
    import torch.distributed as dist
    import pwcnet_b200 as P
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        sys.stdout.flush(); saved = os.dup(1); os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev); dist.barrier(); torch.cuda.synchronize()
        finally:
            sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    B, h, w = args.batch, TRAIN_H, TRAIN_W
    precision = args.precision or P.model.DEFAULT_PRECISION
    model = P.PWCDCNet(weights=P.glorot_init(2), precision=precision, device=dev)
    trainer = P.Trainer(model, lr=1e-4, gamma=4e-4)
    rng = np.random.default_rng(2000 + rank)
    host0 = torch.from_numpy(rng.random((B, h, w, 3), dtype=np.float32)).pin_memory()
    host1 = torch.from_numpy(rng.random((B, h, w, 3), dtype=np.float32)).pin_memory()
    hostg = torch.from_numpy(rng.normal(0, 5, (B, h, w, 2)).astype(np.float32)).pin_memory()
    dev0, dev1, devg = host0.to(dev), host1.to(dev), hostg.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        trainer.step(dev0, dev1, devg)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for s, e in ev:
        flush.fill_(1)
        s.record()
        out = trainer.step(dev0, dev1, devg)
        e.record()
    barrier()
    dev_ms = sum(s.elapsed_time(e) for s, e in ev)
    # e2e: pinned host batch in, loss/EPE scalars back on the host, every step.  TrainStream stages batch i+1 (H2D on a
    # copy stream) while step i computes; sync_value is the un-pipelined Trainer.step(host batch).
    res_host = torch.empty(3, dtype=torch.float32).pin_memory()
    def sync_step():
        loss, lms, epe = trainer.step(host0, host1, hostg)
        res_host.copy_(torch.stack([loss, lms, epe]), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    for _ in range(2):
        sync_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_sync = max(3, args.steps // 4)
    e0.record()
    for _ in range(n_sync):
        sync_step()
    e1.record()
    barrier()
    sync_ms = e0.elapsed_time(e1) / n_sync
    ts = P.TrainStream(trainer, depth=2)
    def piped(n):
        ts.submit(host0, host1, hostg)                     # the first copy is exposed, the others overlap a step
        for k in range(n):
            if k + 1 < n:
                ts.submit(host0, host1, hostg)
            loss, lms, epe = ts.step()
            res_host.copy_(torch.stack([loss, lms, epe]), non_blocking=True)
            torch.cuda.current_stream().synchronize()     # the caller reads the loss of every step
    piped(2)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    piped(args.steps)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches_per_step = trainer.launches_per_step()      # counted (after the timed regions: it re-runs a backward)
    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = t.tolist()
    total_pairs = B * world * args.steps
    if rank == 0:
        cpu_baseline = None
        if not args.no_cpu_baseline:
            pps, cores, _ = cpu_oracle_train_pairs_per_s(3, h, w)
            cpu_baseline = {"value": pps, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": "3 training steps on one 384x1024 pair; restated reference + torch autograd on CPU"}
        line = {"metric": TRAIN_METRIC, "value": total_pairs / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32" if precision == "fp32" else precision,
                "data": "synthetic",
                "config": {"workload": "PWCDCNet(use_dc=False) training step (forward, multiscale L2 loss + 4e-4 l2 regulariser, "
                                       "backward, gradient all-reduce, TF-Adam), batch=%d synthetic 384x1024 pairs + random GT "
                                       "flow per GPU (BASELINE config 5; 436x1024 is not divisible by 64)" % B,
                           "global_batch": B * world, "parallelism": f"dp{world} (one NCCL all-reduce of the 20.1 MB flat gradient per step)",
                           "conv_path": precision + ": tcgen05 forward, dgrad and wgrad (3 x fp16 split, fp32 accumulation); "
                                        "small-channel / stride-2 backward layers on CUDA cores in fp32",
                           "l2": "256 MiB write between timed iterations"},
                "e2e": {"value": total_pairs / (e2e_ms * 1e-3), "unit": UNIT,
                        "h2d_bytes_per_step": int(host0.numel() + host1.numel() + hostg.numel()) * 4, "d2h_bytes_per_step": 12,
                        "sync_value": B * world / (sync_ms * 1e-3),
                        "what": "TrainStream(trainer, depth=2).submit/step: pinned host images + GT flow -> device staging on a copy "
                                "stream (batch i+1 overlaps step i), loss, multiscale loss and EPE copied back and read every step; "
                                "sync_value = Trainer.step(host batch), no overlap"},
                "gpu_launches": args.steps * launches_per_step, "clocks": clocks,
                "loss": float(out[0].item()), "cpu_baseline": cpu_baseline}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _init_dist(dev):
    """NCCL prints its version banner on stdout at communicator creation; keep stdout = the one JSON line."""
    import torch
    import torch.distributed as dist
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


def _bind_to_gpu_numa_node(local: int):
    """Pin this rank's host threads to the CPUs of its GPU's NUMA node, so pinned staging buffers are allocated there
    (first touch) and the copy threads do not cross sockets.  Best effort: silently skipped when sysfs has no answer."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), "pci_bus_id") else None
        out = subprocess.run(["nvidia-smi", f"--id={local}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if not out:
            return None
        bdf = out[-12:] if len(out) >= 12 else out          # 00000000:1B:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            lo, _, hi = part.partition("-")
            ids.update(range(int(lo), int(hi or lo) + 1))
        os.sched_setaffinity(0, ids)
        return node
    except Exception:
        return None


def roofline_cost_volume(P, torch, dev, B, n_sets=4, reps=50):
    """Level-2 cost volume as the model runs it (tcgen05 quadrant-block kernel, split fp16 operands, whole-sector writes
    into the 81+7-word head of the 160-wide estimator concat buffer), timed live: `n_sets` rotating operand / output
    sets (n_sets x 140 MB of touched data > 126 MB L2: every launch misses L2 for all of its inputs, and pays the write-back of its
    predecessor's output), launches back to back, CUDA events around the whole sequence."""
    h2, w2, C = H // 4, W // 4, 32
    g = torch.Generator(device=dev).manual_seed(0)
    sets = []
    for _ in range(n_sets):
        f0 = torch.randn((B, h2, w2, C), device=dev, generator=g)
        f1 = torch.randn((B, h2, w2, C), device=dev, generator=g)
        buf = torch.zeros((B, h2, w2, 160), device=dev)
        sets.append((P.ops.split_f16(f0, scale=1.0 / C), P.ops.split_f16(f1), buf[..., :81]))
    def launch(i):
        a, b, o = sets[i % n_sets]
        P.ops.cost_volume_split(a, b, 0.1, out=o, prescaled=True, slot=True)
    for i in range(10 * n_sets):   # warm-up: tensor maps, clocks, L2 state of the rotation
        launch(i)
    torch.cuda.synchronize()
    n = reps * n_sets
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        launch(i)
    e1.record()
    torch.cuda.synchronize()
    cv_us = 1e3 * e0.elapsed_time(e1) / n
    alg_bytes = 4 * h2 * w2 * (2 * C + 81) * B
    peak, peak_src = _peaks()
    achieved = alg_bytes / (cv_us * 1e-6) / 1e9
    traffic, tsrc = None, None
    tpath = os.path.join(ROOT, "profiles", "r02_cost_volume_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic, tsrc = tj.get("dram_bytes_per_launch"), tj.get("source")
    return {"kernel": "cost_volume_quad_kernel (tcgen05 band GEMM, 3 x fp16 split, fp32 accumulate) level-2 112x256x32 -> 81-ch "
                      "slot of the 160-wide concat buffer, B=%d" % B,
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_source": tsrc, "us_per_launch": cv_us, "algorithmic_bytes": alg_bytes, "peak_source": peak_src,
            "launches_timed": n,
            "method": f"{n_sets} rotating operand/output sets ({n_sets} x 140 MB > L2), {n} back-to-back launches between two CUDA events"}


def roofline_conv(P, torch, dev, B, flush):
    """The kernel that dominates the step by time: the halo-resident tcgen05 conv, level-4 estimator layer 128 -> 128."""
    from pwcnet_b200 import ops_tc
    h2, w2 = H // 4, W // 4
    g = torch.Generator(device=dev).manual_seed(1)
    xc = torch.randn((B, h2, w2, 128), device=dev, generator=g)
    kc = torch.randn((3, 3, 128, 128), device=dev, generator=g) / 34.0
    bc = torch.zeros(128, device=dev)
    yc = torch.empty((B, h2, w2, 128), device=dev)
    wpk = ops_tc.pack_weights_f16(kc)
    for _ in range(3):
        ops_tc.conv3x3_tc_f16(xc, wpk, bc, 128, 128, alpha=0.1, out=yc)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    for s_, e_ in ev:
        flush.fill_(3)
        s_.record()
        ops_tc.conv3x3_tc_f16(xc, wpk, bc, 128, 128, alpha=0.1, out=yc)
        e_.record()
    torch.cuda.synchronize()
    conv_us = 1e3 * float(np.mean([s_.elapsed_time(e_) for s_, e_ in ev]))
    conv_flops = 2.0 * 9 * 128 * 128 * B * h2 * w2
    tpeak, tpeak_src = _tensor_peak()
    conv_tflops = conv_flops / (conv_us * 1e-6) / 1e12
    return {"kernel": "conv3x3_tc_halo_kernel, level-4 estimator conv 128->128 at 112x256, B=%d (3 x fp16 split: every useful "
                      "flop costs 3 fp16 MMA flops)" % B,
            "bound": "tensor", "achieved": conv_tflops, "peak": tpeak, "unit": "TFLOP/s", "frac": conv_tflops / tpeak,
            "mma_issue_tflops": 3 * conv_tflops, "mma_issue_frac": 3 * conv_tflops / tpeak, "traffic": None,
            "us_per_launch": conv_us, "algorithmic_flops": conv_flops, "peak_source": tpeak_src}


def train_object(P, torch, dist, dev, rank, world, B, steps, warmup, precision):
    """BASELINE config 5 inside the default line: training step at 384x1024, B pairs per GPU, NCCL all-reduce of the flat
    gradient at world > 1.  Device-timed like `value`; `allreduce_exposed_ms` = step time minus the same step with the
    collective skipped."""
    h, w = TRAIN_H, TRAIN_W
    model = P.PWCDCNet(weights=P.glorot_init(2), precision=precision, device=dev, cv_pipeline="default")
    trainer = P.Trainer(model, lr=1e-4, gamma=4e-4)
    rng = np.random.default_rng(2000 + rank)
    d0 = torch.from_numpy(rng.integers(0, 256, (B, h, w, 3), dtype=np.uint8)).to(dev)
    d1 = torch.from_numpy(rng.integers(0, 256, (B, h, w, 3), dtype=np.uint8)).to(dev)
    dg = torch.from_numpy(rng.normal(0, 5, (B, h, w, 2)).astype(np.float32)).to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = trainer.step(d0, d1, dg)
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / n, out

    for _ in range(warmup):
        trainer.step(d0, d1, dg)
    ms, out = timed(steps)
    ms_noar = None
    if world > 1:
        trainer.skip_allreduce = True          # measurement only: same step without the collective
        timed(2)
        ms_noar, _ = timed(steps)
        trainer.skip_allreduce = False
    t = torch.tensor([ms, ms_noar if ms_noar is not None else ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_noar = t.tolist()
    return {"metric": TRAIN_METRIC, "value": B * world / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
            "batch_per_gpu": B, "global_batch": B * world, "ranks": world,
            "allreduce": "bucketed NCCL SUM all-reduce of the flat fp32 gradient (20.1 MB) on a side stream, overlapped with "
                         "the pyramid backward" if world > 1 else "none (1 rank)",
            "allreduce_exposed_ms": (ms - ms_noar) if world > 1 else 0.0,
            "loss": float(out[0].item()), "gpu_launches_per_step": trainer.launches_per_step(),
            "workload": "PWCDCNet(use_dc=False) training step (forward, multiscale L2 loss + 4e-4 l2 regulariser, backward, gradient "
                        "all-reduce, TF-Adam), synthetic uint8 384x1024 pairs + random GT flow (BASELINE config 5; 436x1024 is "
                        "not divisible by 64)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="infer = the headline metric (default; carries a `train` object); train = BASELINE config 5 alone")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8, help="pairs per GPU per step")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--precision", default=None, choices=["fp32", "3xf16", "3xtf32", "tf32", "cudnn"])
    ap.add_argument("--cpu-pairs", type=int, default=20, help="pairs timed for cpu_baseline (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the `train` object of the default line")
    ap.add_argument("--min-seconds", type=float, default=2.0,
                    help="each of the K timed steps repeats the batch so that the timed region lasts at least this long")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.mode == "train":
        return run_train_reference(args) if args.impl == "reference" else run_train(args)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import pwcnet_b200 as P

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU fallback")
    numa = _bind_to_gpu_numa_node(local)          # before any pinned allocation
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        _init_dist(dev)
    if args.gpus != world and rank == 0 and world > 1:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    B = args.batch
    precision = args.precision or P.model.DEFAULT_PRECISION
    model = P.PWCDCNet(weights=P.glorot_init(2), precision=precision, device=dev)
    rng = np.random.default_rng(1000 + rank)
    # the reference's inputs are uint8 images divided by 255 on the host (test.py:31-33, train.py:122): the host API takes
    # the bytes and does the division on the device (bit-identical), so a quarter of the bytes cross PCIe
    host0 = torch.from_numpy(rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)).pin_memory()
    host1 = torch.from_numpy(rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)).pin_memory()
    dev0, dev1 = host0.to(dev), host1.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- roofline kernels first: timed alone on a cool GPU (the
    # long timed regions below run into the 1 kW power cap; MEASURED_PEAKS' hbm_gbs is a burst figure as well)
    roofline = roof_conv = None
    if rank == 0:
        roofline = roofline_cost_volume(P, torch, dev, B)
        roof_conv = roofline_conv(P, torch, dev, B, flush)
    barrier()

    # ---------------------------------------------------------------- N-GPU == 1-GPU evidence: every rank runs the same
    # seed-0 probe pair and reports a checksum of its flow (bit-for-bit comparable across ranks and across --gpus runs)
    pr = np.random.default_rng(0)
    p0 = pr.integers(0, 256, (1, H, W, 3), dtype=np.uint8)
    p1 = np.roll(p0, (3, -5), axis=(1, 2))
    pf, _ = model(p0, p1)
    import hashlib
    digest = hashlib.sha256(pf.cpu().numpy().tobytes()).hexdigest()[:16]
    csum = torch.tensor([int(digest, 16) % (1 << 52)], dtype=torch.float64, device=dev)
    all_cs = [csum.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(all_cs, csum)
    probe_equal = all(float(c.item()) == float(csum.item()) for c in all_cs)

    # ---------------------------------------------------------------- value: inputs resident in HBM
    for _ in range(args.warmup):
        model(dev0, dev1)
    barrier()
    # inner repeats so that the timed region is >= --min-seconds (thermal / power steady state, not a 70 ms burst)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); model(dev0, dev1); model(dev0, dev1); e1.record(); torch.cuda.synchronize()
    est_ms = e0.elapsed_time(e1) / 2
    inner = max(1, int(np.ceil(args.min_seconds * 1e3 / (est_ms * args.steps))))
    if world > 1:
        ti = torch.tensor([inner], dtype=torch.int64, device=dev)
        dist.all_reduce(ti, op=dist.ReduceOp.MAX)
        inner = int(ti.item())
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps * inner)]
    for s, e in ev:
        flush.fill_(1)                       # evict L2 between timed iterations (outside the event pair)
        s.record()
        model(dev0, dev1)
        e.record()
    barrier()
    times = [s.elapsed_time(e) for s, e in ev]
    dev_ms = sum(times)
    burst_ms = sum(times[:args.steps])       # the first K passes: what a K-step run without repeats would report

    # ---------------------------------------------------------------- e2e: host buffers through the public API
    # (a) synchronous reference-shaped call: model(host uint8 images) then copy the results back, one request at a time
    ff, pyr = model(host0, host1)
    outs_dev = [ff] + list(pyr)
    out_host = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in outs_dev]   # pinned result buffers

    def e2e_sync_step():
        ff, pyr = model(host0, host1)                        # H2D of both images from pinned memory inside
        for h, d in zip(out_host, [ff] + list(pyr)):         # D2H of the final flow + the 5 pyramid flows
            h.copy_(d, non_blocking=True)
        torch.cuda.current_stream().synchronize()            # results are on the host when the step ends

    for _ in range(2):
        e2e_sync_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        e2e_sync_step()
    e1.record()
    barrier()
    e2e_sync_ms = e0.elapsed_time(e1)

    # (b) pipelined public API (InferenceStream, 2 requests in flight): every request still starts in pinned
    #     host memory and ends in pinned host memory; H2D / D2H of neighbouring requests overlap the forward
    stream = P.InferenceStream(model, depth=2)
    for _ in range(2):
        stream.collect(stream.submit(host0, host1))
    barrier()
    n_e2e = args.steps * max(1, inner // 4)
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    prev = None
    for _ in range(n_e2e):
        tk = stream.submit(host0, host1)
        if prev is not None:
            stream.collect(prev)
        prev = tk
    res = stream.collect(prev)
    out_host = [res[0]] + list(res[1])
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    wall_e2e = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    h2d = int(host0.numel() * host0.element_size() + host1.numel() * host1.element_size())
    d2h = int(sum(t.numel() * t.element_size() for t in out_host))

    t = torch.tensor([dev_ms, e2e_ms / n_e2e, e2e_sync_ms, burst_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_step_ms, e2e_sync_ms, burst_ms = t.tolist()
    n_timed = args.steps * inner
    value = B * world * n_timed / (dev_ms * 1e-3)
    e2e_val = B * world / (e2e_step_ms * 1e-3)

    # ---------------------------------------------------------------- training (BASELINE config 5) in the same line
    train = None
    if not args.no_train and precision != "cudnn":
        del stream
        train = train_object(P, torch, dist, dev, rank, world, B, max(5, min(args.steps, 20)), 3, precision)

    if rank == 0:
        cpu_baseline = None
        if not args.no_cpu_baseline:
            pps, cores, _ = cpu_oracle_pairs_per_s(args.cpu_pairs)
            cpu_baseline = {"value": pps, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"{args.cpu_pairs} forward passes of one 448x1024 pair (same net, same glorot "
                                      "weights); restated reference on torch-CPU, not TensorFlow 1.8"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / n_timed, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if precision in ("fp32", "cudnn") else precision, "data": "synthetic",
            "config": {"workload": "PWCDCNet(use_dc=False) inference, batch=%d synthetic 448x1024 pairs per GPU "
                                   "(BASELINE config 2 shape), random-init glorot weights" % B,
                       "global_batch": B * world, "parallelism": f"dp{world} (no data-path collective)",
                       "conv_path": precision, "cuda_graph": True, "inner_repeats": inner,
                       "timed_region": f"each of the {args.steps} timed steps runs {inner} passes over the batch ({n_timed} passes, "
                                       f"{dev_ms / 1e3:.2f} s of device time); value and ms_per_step are per pass",
                       "l2": "256 MiB write between timed passes; per-pass CUDA-event intervals summed"},
            "burst_value": B * world * args.steps / (burst_ms * 1e-3),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_step_ms, "wall_ms_per_step": 1e3 * wall_e2e / n_e2e, "steps_timed": n_e2e,
                    "h2d_gbs_per_rank": h2d / (e2e_step_ms * 1e-3) / 1e9, "d2h_gbs_per_rank": d2h / (e2e_step_ms * 1e-3) / 1e9,
                    "numa_node_bound": numa,
                    "sync_value": B * world * args.steps / (e2e_sync_ms * 1e-3),
                    "what": "InferenceStream(model, depth=2).submit/collect: pinned host uint8 images (the reference's input before "
                            "its /255.0) -> flows_final + 5 pyramid flows (float32) in pinned host memory, H2D/D2H of neighbouring "
                            "requests overlapped with the forward; sync_value = one PWCDCNet.__call__(host uint8) at a time + D2H, no overlap"},
            "gpu_launches": n_timed * (model.launches_per_forward() + 0),
            "probe": {"sha256_16": digest, "equal_across_ranks": probe_equal,
                      "what": "flows_final of a fixed seed-0 uint8 pair (1x448x1024), identical on every rank and for every --gpus"},
            "clocks": clocks, "roofline": roofline, "roofline_conv": roof_conv, "cpu_baseline": cpu_baseline, "train": train,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
