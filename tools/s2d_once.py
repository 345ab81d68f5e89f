"""One stride-2 (space-to-depth) halo conv followed by a stride-1 conv: python tools/s2d_once.py C Co B H W [split]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pwcnet_b200 as P
from pwcnet_b200 import ops_tc
C, Co, B, H, W = (int(a) for a in sys.argv[1:6])
split = len(sys.argv) > 6
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((B, H, W, C), device="cuda", generator=g)
k = torch.randn((3, 3, C, Co), device="cuda", generator=g) * 0.1
k2 = torch.randn((3, 3, Co, Co), device="cuda", generator=g) * 0.1
b = torch.randn((Co,), device="cuda", generator=g) * 0.1
wp = ops_tc.pack_weights_f16(ops_tc.s2d_reindex(k))
wp2 = ops_tc.pack_weights_f16(k2)
for _ in range(3):
    if split:
        ys = torch.empty((B, H // 2, W // 2, 2 * Co), dtype=torch.float16, device="cuda")
        ops_tc.conv3x3_s2_tc_f16(x, wp, b, C, Co, alpha=0.1, out_split=ys)
        z = ops_tc.conv3x3_tc_f16_split(ys, wp2, b, Co, Co, alpha=0.1)
    else:
        y = ops_tc.conv3x3_s2_tc_f16(x, wp, b, C, Co, alpha=0.1)
        z = ops_tc.conv3x3_tc_f16(y, wp2, b, Co, Co, alpha=0.1)
torch.cuda.synchronize()
ref = P.ops.conv3x3(P.ops.conv3x3(x, k, b, stride=2, alpha=0.1), k2, b, alpha=0.1)
print("max err", float((z - ref).abs().max()))
