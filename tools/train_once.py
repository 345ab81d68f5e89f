"""Two training steps at BASELINE config 5 shape (target for ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pwcnet_b200 as P
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
model = P.PWCDCNet(weights=P.glorot_init(2))
tr = P.Trainer(model)
rng = np.random.default_rng(0)
i0 = torch.from_numpy(rng.random((B, 384, 1024, 3), dtype=np.float32)).cuda()
i1 = torch.from_numpy(rng.random((B, 384, 1024, 3), dtype=np.float32)).cuda()
gt = torch.from_numpy(rng.normal(0, 5, (B, 384, 1024, 2)).astype(np.float32)).cuda()
for _ in range(steps):
    loss, lms, epe = tr.step(i0, i1, gt)
torch.cuda.synchronize()
print("loss", loss.item(), "epe", epe.item())
