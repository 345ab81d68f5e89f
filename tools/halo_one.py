"""One level-4 estimator conv (128->128, B=8, 112x256) a few times: ncu target for the halo-resident kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pwcnet_b200 import ops_tc, ops_bwd
x = torch.randn(8, 112, 256, 128, device="cuda"); k = torch.randn(3, 3, 128, 128, device="cuda") / 34; b = torch.zeros(128, device="cuda")
y = torch.empty(8, 112, 256, 128, device="cuda"); wp = ops_tc.pack_weights_f16(k)
for _ in range(4): ops_tc.conv3x3_tc_f16(x, wp, b, 128, 128, alpha=0.1, out=y)
dy = torch.randn(8, 112, 256, 128, device="cuda") * 1e-3; dw = torch.zeros(3, 3, 128, 128, device="cuda")
xT = ops_bwd.tsplit(x, conv_input=True); dyT = ops_bwd.tsplit(dy)
for _ in range(4): ops_bwd.conv3x3_wgrad_tc(xT, dyT, dw, (8, 112, 256, 128), 128)
torch.cuda.synchronize()
