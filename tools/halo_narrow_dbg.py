"""Per-CTA timeline (PWC_HALO_DEBUG=1) of the halo conv on one narrow layer: python tools/halo_narrow_dbg.py C Co B H W"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pwcnet_b200 as P
from pwcnet_b200 import ops_tc
C, Co, B, H, W = (int(a) for a in sys.argv[1:6])
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((B, H, W, C), device="cuda", generator=g)
k = torch.randn((3, 3, C, Co), device="cuda", generator=g) * 0.1
b = torch.randn((Co,), device="cuda", generator=g) * 0.1
wp = ops_tc.pack_weights_f16(k)
for _ in range(3):
    ops_tc.conv3x3_tc_f16(x, wp, b, C, Co, alpha=0.1)      # warm-up: the timeline below is of a warm launch
torch.cuda.synchronize()
os.environ["PWC_HALO_DEBUG"] = "1"
ops_tc.conv3x3_tc_f16(x, wp, b, C, Co, alpha=0.1)
torch.cuda.synchronize()
