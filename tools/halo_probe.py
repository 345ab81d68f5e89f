"""Halo-resident f16 conv (conv_tc_halo.cu) vs the streaming kernel and fp64: accuracy and timing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pwcnet_b200 as P
from pwcnet_b200 import ops_tc
torch.manual_seed(0)
def ref64(x, k, b, d=1):
    x = x.double().cpu(); k = k.double().cpu()
    y = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), k.permute(3, 2, 0, 1), b.double().cpu(), padding=d, dilation=d)
    return y.permute(0, 2, 3, 1)
def run(x, wp, b, Cin, Cout, halo, out=None, d=1):
    os.environ["PWC_CONV_HALO"] = "1" if halo else "0"
    return ops_tc.conv3x3_tc_f16(x, wp, b, Cin, Cout, dilation=d, alpha=0.1, out=out)
def case(B, H, W, Cin, Cout, cs=None, d=1):
    cs = cs or Cin
    x = torch.randn(B, H, W, cs, device="cuda")[..., :Cin]
    k = torch.randn(3, 3, Cin, Cout, device="cuda") / np.sqrt(9 * Cin); b = torch.randn(Cout, device="cuda") * 0.1
    r = ref64(x, k, b, d); r = torch.maximum(0.1 * r, r)
    wp = ops_tc.pack_weights_f16(k)
    yh = run(x, wp, b, Cin, Cout, True, d=d); torch.cuda.synchronize()
    ys = run(x, wp, b, Cin, Cout, False, d=d); torch.cuda.synchronize()
    eh = (yh.double().cpu() - r).abs().max().item(); es = (ys.double().cpu() - r).abs().max().item()
    print(f"B{B} {H}x{W} Cin{Cin}(cs{cs}) Cout{Cout} d{d}: err halo {eh:.2e}  streaming {es:.2e}  max|y| {r.abs().max().item():.1f}", flush=True)
def timeit(B, H, W, Cin, Cout, cs=None, iters=10, d=1):
    cs = cs or Cin
    x = torch.randn(B, H, W, cs, device="cuda")[..., :Cin]
    k = torch.randn(3, 3, Cin, Cout, device="cuda") / np.sqrt(9 * Cin); b = torch.zeros(Cout, device="cuda"); y = torch.empty(B, H, W, Cout, device="cuda")
    wp = ops_tc.pack_weights_f16(k)
    for halo in (True, False):
        for _ in range(3): run(x, wp, b, Cin, Cout, halo, y, d=d)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters): run(x, wp, b, Cin, Cout, halo, y, d=d)
        e.record(); torch.cuda.synchronize()
        us = s.elapsed_time(e) * 1e3 / iters
        print(f"time B{B} {H}x{W} {Cin}->{Cout} d{d} {'halo' if halo else 'stream'}: {us:.1f} us  {2*9*Cin*Cout*B*H*W/us/1e6:.1f} TFLOP/s", flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "time":
    timeit(8, 112, 256, 128, 128); timeit(8, 112, 256, 147, 128, cs=148); timeit(8, 112, 256, 64, 32); timeit(8, 112, 256, 96, 64)
    timeit(16, 112, 256, 32, 32); timeit(16, 224, 512, 16, 16); timeit(8, 56, 128, 128, 128)
    timeit(8, 112, 256, 128, 128, d=2); timeit(8, 112, 256, 128, 128, d=4); timeit(8, 112, 256, 128, 96, d=8); timeit(8, 112, 256, 96, 64, d=16)
else:
    case(1, 4, 128, 32, 32); case(1, 9, 130, 32, 16); case(2, 14, 256, 128, 128); case(1, 7, 200, 147, 128, cs=148)
    case(1, 5, 96, 64, 64); case(2, 6, 300, 16, 16); case(1, 3, 512, 96, 48); case(1, 2, 128, 34, 128, cs=36)
    case(1, 9, 130, 32, 32, d=2); case(2, 20, 256, 128, 128, d=4); case(1, 30, 200, 128, 96, d=8); case(1, 40, 256, 96, 64, d=16); case(1, 5, 100, 64, 32, d=16)
