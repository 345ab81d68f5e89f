import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pwcnet_b200 as P
from pwcnet_b200 import ops_tc
ns = int(sys.argv[1]); Cin = int(sys.argv[2]); Cout = int(sys.argv[3])
B,H,W = 8,112,256
x = torch.randn(B,H,W,Cin, device="cuda")
k = torch.randn(3,3,Cin,Cout, device="cuda") / np.sqrt(9*Cin); b = torch.zeros(Cout, device="cuda")
wp = ops_tc.pack_weights(k); y = torch.empty(B,H,W,Cout, device="cuda")
for _ in range(4):
    ops_tc.conv3x3_tc(x, wp, b, Cin, Cout, alpha=0.1, n_split=ns, out=y)
torch.cuda.synchronize()
