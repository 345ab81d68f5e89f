"""Level-1/2 pyramid wgrad shapes (small-channel CUDA-core kernel) + level-2 cost-volume backward: targets for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pwcnet_b200 as P
from pwcnet_b200 import ops_bwd
torch.manual_seed(0)
def t(*s): return torch.randn(*s, device="cuda")
cases = [(16, 192, 512, 16, 16, 1), (16, 96, 256, 32, 32, 1), (16, 384, 1024, 3, 16, 2)]
for (B, H, W, ci, co, s) in cases:
    x, dy = t(B, H, W, ci), t(B, H // s, W // s, co)
    dw, db = torch.zeros(3, 3, ci, co, device="cuda"), torch.zeros(co, device="cuda")
    for _ in range(2):
        ops_bwd.conv3x3_wgrad(x, dy, dw, db, stride=s)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5):
        ops_bwd.conv3x3_wgrad(x, dy, dw, db, stride=s)
    e1.record(); torch.cuda.synchronize()
    print("wgrad_small", (B, H, W, ci, co, s), round(e0.elapsed_time(e1) / 5 * 1000, 1), "us")
B, H, W, C = 8, 96, 256, 32
f0, f1, g, cv = t(B, H, W, C), t(B, H, W, C), t(B, H, W, 81), t(B, H, W, 81)
df0, df1 = torch.zeros_like(f0), torch.zeros_like(f1)
for _ in range(2):
    ops_bwd.cost_volume_bwd(g, cv, f0, f1, df0, df1)
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(5):
    ops_bwd.cost_volume_bwd(g, cv, f0, f1, df0, df1)
e1.record(); torch.cuda.synchronize()
print("cost_volume_bwd", (B, H, W, C), round(e0.elapsed_time(e1) / 5 * 1000, 1), "us")
