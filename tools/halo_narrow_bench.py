"""Halo-conv layers at the network's shapes: timing + parity vs the exact CUDA-core conv.
Env: PWC_HALO_EPI_DIRECT=1 (round-1 epilogue: 16-byte stores at the pixel stride), PWC_HALO_EXP=2 (no epilogue stores),
PWC_HALO_SETS=n (rotating accumulator sets)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pwcnet_b200 as P
from pwcnet_b200 import ops_tc
g = torch.Generator(device="cuda").manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("PWC_HALO")) or "default"
for (B, H, W, C, Co) in ((16, 224, 512, 16, 16), (16, 112, 256, 32, 32), (8, 112, 256, 32, 16), (16, 56, 128, 64, 64),
                         (8, 112, 256, 128, 128), (8, 112, 256, 128, 96), (8, 112, 256, 96, 64), (8, 112, 256, 64, 32),
                         (8, 28, 64, 32, 64)):
    x = torch.randn((B, H, W, C), device="cuda", generator=g)
    k = torch.randn((3, 3, C, Co), device="cuda", generator=g) * 0.1
    b = torch.randn((Co,), device="cuda", generator=g) * 0.1
    wp = ops_tc.pack_weights_f16(k)
    y = ops_tc.conv3x3_tc_f16(x, wp, b, C, Co, alpha=0.1)
    ref = P.ops.conv3x3(x, k, b, alpha=0.1)
    err = float((y - ref).abs().max())
    ts = []
    for _ in range(10):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ops_tc.conv3x3_tc_f16(x, wp, b, C, Co, alpha=0.1, out=y); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    print(f"[{tag}] {C}->{Co} at {B}x{H}x{W}: {sum(ts)/len(ts):.1f} us (min {min(ts):.1f}), max-abs err vs fp32 conv {err:.2e}")
