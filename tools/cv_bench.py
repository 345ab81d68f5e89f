"""Launches the level-2 cost-volume kernel (B x 112 x 256 x 32) a few times: target for ncu captures
and quick timing.  usage: python tools/cv_bench.py [B] [iters] [fused]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pwcnet_b200 as P

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
fused = len(sys.argv) > 3 and sys.argv[3] == "fused"
slot = len(sys.argv) > 3 and sys.argv[3] == "slot"   # write into the 148-wide concat buffer of level 2, as the model does
h, w, C = 112, 256, 32
g = torch.Generator(device="cuda").manual_seed(0)
f0 = torch.randn((B, h, w, C), device="cuda", generator=g)
f1 = torch.randn((B, h, w, C), device="cuda", generator=g)
flow = torch.randn((B, h, w, 2), device="cuda", generator=g) * 3
cv = torch.empty((B, h, w, 148), device="cuda")[..., :81] if slot else torch.empty((B, h, w, 81), device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
# PWC_FLUSH=clean: after the write-flush, READ a second 256 MiB buffer so that L2 holds only clean lines when the timed
# kernel starts (a write-only flush leaves ~126 MB of dirty lines whose write-back is charged to the timed kernel)
clean = os.environ.get("PWC_FLUSH") == "clean"
wide = os.environ.get("PWC_WIDE", "0") != "0"     # whole-sector slot writes (pwc_cost_volume_split_slot_fwd)
flush2 = torch.ones(64 << 20, dtype=torch.float32, device="cuda") if clean else None
split = len(sys.argv) > 3 and sys.argv[3].startswith("split")
if split:
    if sys.argv[3].startswith("splitslot"):           # splitslot[N]: 81-channel slot of an N-wide buffer (default 148)
        pitch = int(sys.argv[3][9:] or 148)
        cv = torch.empty((B, h, w, pitch), device="cuda")[..., :81]
    f0s, f1s = P.ops.split_f16(f0, scale=1.0 / C), P.ops.split_f16(f1)
def run():
    if split:
        P.ops.cost_volume_split(f0s, f1s, out=cv, prescaled=True, slot=wide)
    elif fused:
        P.ops.warp_cost_volume(f0, f1, flow, 5.0, out=cv)
    else:
        P.ops.cost_volume(f0, f1, out=cv)
for _ in range(3):
    run()
if os.environ.get("PWC_ROTATE"):
    # rotating operand/output sets: every launch misses L2 for all inputs (sets x 133 MB > 126 MB L2), launches back to back,
    # per-launch time = event interval / launches (no flush kernel, no per-launch event gap)
    nset = int(os.environ["PWC_ROTATE"])
    sets = []
    for i in range(nset):
        a = torch.randn((B, h, w, C), device="cuda", generator=g); b_ = torch.randn((B, h, w, C), device="cuda", generator=g)
        o = torch.empty_like(cv.as_strided(cv.size(), cv.stride()).contiguous()) if False else torch.empty((B, h, w, cv.stride(2)), device="cuda")[..., :81]
        sets.append((P.ops.split_f16(a, scale=1.0 / C), P.ops.split_f16(b_), o) if split else (a, b_, o))
    def run_set(i):
        x, y, o = sets[i % nset]
        if split:
            P.ops.cost_volume_split(x, y, out=o, prescaled=True, slot=wide)
        else:
            P.ops.cost_volume(x, y, out=o)
    for i in range(nset):
        run_set(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = iters * nset
    s.record()
    for i in range(n):
        run_set(i)
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) * 1e3 / n
    alg = 4 * h * w * (2 * C + 81) * B
    print(f"cost_volume level-2 B={B} mode={sys.argv[3] if len(sys.argv) > 3 else 'dense'} split={os.environ.get('PWC_CV_SPLIT', '-')} rotate={nset} x {iters}: {us:.1f} us/launch, {alg/us/1e3:.0f} GB/s algorithmic, frac of 6550 = {alg/us/1e3/6550:.3f}")
    sys.exit(0)
ts = []
for _ in range(iters):
    flush.fill_(1)
    if clean:
        flush2.sum()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); run(); e.record()
    torch.cuda.synchronize()
    ts.append(s.elapsed_time(e) * 1e3)
alg = 4 * h * w * (2 * C + 81 + (2 if fused else 0)) * B
import statistics
us = statistics.mean(ts)
print(f"cost_volume level-2 B={B} mode={sys.argv[3] if len(sys.argv) > 3 else 'dense'} split={os.environ.get('PWC_CV_SPLIT', '-')} flush={'clean' if clean else 'dirty'}: {us:.1f} us/launch (min {min(ts):.1f}), {alg/us/1e3:.0f} GB/s algorithmic, frac of 6550 = {alg/us/1e3/6550:.3f}")
