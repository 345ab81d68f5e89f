import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pwcnet_b200 as P
from pwcnet_b200 import ops_bwd
B, H, W, Cin, Cout = 2, 9, 13, 32, 64
x = torch.randn(B, H, W, Cin, device="cuda"); dy = torch.randn(B, H, W, Cout, device="cuda") * 1e-3
db = torch.zeros(Cout, device="cuda")
xT = ops_bwd.tsplit(x, conv_input=True); torch.cuda.synchronize(); print("tsplit x ok")
dyT = ops_bwd.tsplit(dy, db=db); torch.cuda.synchronize(); print("tsplit dy ok", float((db - dy.sum((0,1,2))).abs().max()))
Wp = 16
pl = B*Cin*H*Wp; h = xT[2*pl:3*pl].view(B, Cin, H, Wp).float(); l = xT[3*pl:4*pl].view(B, Cin, H, Wp).float()   # copy kx = 1 (no shift)
rec = (h + l / 2048)[..., :W].permute(0, 2, 3, 1)
print("tsplit reconstruct err", float((rec - x).abs().max()))
dw = torch.zeros(3, 3, Cin, Cout, device="cuda")
ops_bwd.conv3x3_wgrad_tc(xT, dyT, dw, (B, H, W, Cin), Cout); torch.cuda.synchronize(); print("wgrad_tc ok")
ref = torch.zeros_like(dw)
ops_bwd.conv3x3_wgrad(x, dy, ref); torch.cuda.synchronize()
print("err", float((dw - ref).abs().max()), "max", float(ref.abs().max()))
