"""Runs PWCDCNet forward (B pairs 448x1024) a few times without CUDA graph: target for ncu launch lists."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pwcnet_b200 as P
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prec = sys.argv[3] if len(sys.argv) > 3 else None
model = P.PWCDCNet(weights=P.glorot_init(2), precision=prec, use_cuda_graph=False)
rng = np.random.default_rng(0)
u8 = os.environ.get("PWC_FWD_F32") is None          # default: uint8 images, as bench.py feeds them
if u8:
    a = torch.from_numpy(rng.integers(0, 256, (B, 448, 1024, 3), dtype=np.uint8)).cuda()
    b = torch.from_numpy(rng.integers(0, 256, (B, 448, 1024, 3), dtype=np.uint8)).cuda()
else:
    a = torch.from_numpy(rng.random((B, 448, 1024, 3), dtype=np.float32)).cuda()
    b = torch.from_numpy(rng.random((B, 448, 1024, 3), dtype=np.float32)).cuda()
for _ in range(n):
    model(a, b)
torch.cuda.synchronize()
print("launches per forward:", model.launches_per_forward())
