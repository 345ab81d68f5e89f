import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pwcnet_b200 import ops_tc
B, H, W, Cin, Cout = [int(a) for a in sys.argv[1:6]]
x = torch.randn(B, H, W, Cin, device="cuda"); k = torch.randn(3, 3, Cin, Cout, device="cuda") / 20; b = torch.zeros(Cout, device="cuda")
y = torch.empty(B, H, W, Cout, device="cuda"); wp = ops_tc.pack_weights_f16(k)
ops_tc.conv3x3_tc_f16(x, wp, b, Cin, Cout, alpha=0.1, out=y); torch.cuda.synchronize()
