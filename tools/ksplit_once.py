"""Level-6 / level-5 pyramid convs on the streaming tcgen05 kernel, timed alone (A/B for PWC_TC_NO_KSPLIT):
python tools/ksplit_once.py"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pwcnet_b200 import ops_tc
torch.cuda.set_device(0)
g = torch.Generator(device="cuda").manual_seed(0)
cases = [("L6 128->192 s2 @14x32x16", 16, 14, 32, 128, 192, 2), ("L6 192->192 @7x16x16", 16, 7, 16, 192, 192, 1),
         ("L5 96->128 s2 @28x64x16", 16, 28, 64, 96, 128, 2), ("192->192 @7x16x64 (B=32)", 64, 7, 16, 192, 192, 1)]
for name, B, H, W, ci, co, stride in cases:
    x = torch.randn((B, H, W, ci), device="cuda", generator=g)
    k = torch.randn((3, 3, ci, co), device="cuda", generator=g) / (3 * ci ** 0.5)
    b = torch.zeros(co, device="cuda")
    wp = ops_tc.pack_weights_f16(k)
    y = ops_tc.conv3x3_tc_f16(x, wp, b, ci, co, alpha=0.1, stride=stride)
    for _ in range(5):
        ops_tc.conv3x3_tc_f16(x, wp, b, ci, co, alpha=0.1, stride=stride, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200
    e0.record()
    for _ in range(n):
        ops_tc.conv3x3_tc_f16(x, wp, b, ci, co, alpha=0.1, stride=stride, out=y)
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"case": name, "us": 1e3 * e0.elapsed_time(e1) / n, "checksum": float(y.double().abs().sum())}))
