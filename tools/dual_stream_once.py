"""Experiment: two forwards in flight on two streams (two models = two plans / CUDA graphs, same weights) vs one after the
other: does the second request fill the SMs the persistent kernels of the first leave idle (tail rounds, coarse levels)?
python tools/dual_stream_once.py [B] [passes]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pwcnet_b200 as P
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = int(sys.argv[2]) if len(sys.argv) > 2 else 200
H, W = 448, 1024
torch.cuda.set_device(0); dev = torch.device("cuda:0")
rng = np.random.default_rng(1)
ims = [torch.from_numpy(rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)).to(dev) for _ in range(4)]
m = [P.PWCDCNet(weights=P.glorot_init(2), device=dev) for _ in range(2)]
ref = []
for k in range(2):
    for _ in range(3):
        ff, _ = m[k](ims[2 * k], ims[2 * k + 1])
    torch.cuda.synchronize()
    ref.append(ff.clone())
def timed(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(n); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
def seq(n):
    for i in range(n):
        m[i & 1](ims[2 * (i & 1)], ims[2 * (i & 1) + 1])
s = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
def dual(n):
    cur = torch.cuda.current_stream()
    for k in range(2):
        s[k].wait_stream(cur)
    for i in range(n):
        with torch.cuda.stream(s[i & 1]):
            m[i & 1](ims[2 * (i & 1)], ims[2 * (i & 1) + 1])
    for k in range(2):
        cur.wait_stream(s[k])
for name, fn in (("sequential", seq), ("dual", dual), ("sequential", seq), ("dual", dual)):
    fn(10)
    ms = timed(fn, N)
    print(json.dumps({"mode": name, "ms_per_pass": ms / N, "pairs_per_s": B * N / (ms * 1e-3)}))
torch.cuda.synchronize()
for k in range(2):
    ff, _ = m[k](ims[2 * k], ims[2 * k + 1]); torch.cuda.synchronize()
    print("bit-identical after dual runs:", bool(torch.equal(ff, ref[k])))
