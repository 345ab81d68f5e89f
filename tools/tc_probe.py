"""First contact with the tcgen05 conv kernel: correctness vs the fp32 direct kernel and a float64
CPU reference, whether kind::tf32 truncates or rounds raw fp32 operands, and timing of the big layers."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pwcnet_b200 as P
from pwcnet_b200 import ops_tc

torch.manual_seed(0)
def ref64(x, k, b, dil, trunc_x=False, round_x=False, trunc_w=False):
    x = x.double().cpu(); k = k.double().cpu()
    def tr(t):
        return (t.float().view(torch.int32) & -8192).view(torch.float32).double()
    def rn(t):
        i = t.float().view(torch.int32)
        return ((i + 0x1000) & -8192).view(torch.float32).double()
    if trunc_x: x = tr(x)
    if round_x: x = rn(x)
    if trunc_w: k = tr(k)
    y = torch.nn.functional.conv2d(x.permute(0,3,1,2), k.permute(3,2,0,1), b.double().cpu(), padding=dil, dilation=dil)
    return y.permute(0,2,3,1)

def case(B,H,W,Cin,Cout,dil,cs=None,alpha=0.1):
    cs = cs or Cin
    buf = torch.randn(B,H,W,cs, device="cuda")
    x = buf[..., :Cin]
    k = torch.randn(3,3,Cin,Cout, device="cuda") / np.sqrt(9*Cin)
    b = torch.randn(Cout, device="cuda")*0.1
    wp = ops_tc.pack_weights(k)
    r = ref64(x,k,b,dil); r = torch.maximum(alpha*r, r)
    out = {}
    for ns in (1,3):
        y = ops_tc.conv3x3_tc(x, wp, b, Cin, Cout, dilation=dil, alpha=alpha, n_split=ns)
        torch.cuda.synchronize()
        out[ns] = (y.double().cpu()-r).abs().max().item()
    yd = P.ops.conv3x3(x, k, b, dilation=dil, alpha=alpha)
    ed = (yd.double().cpu()-r).abs().max().item()
    # truncation vs rounding hypothesis for the 1x path (alpha=1 to keep it linear)
    y1 = ops_tc.conv3x3_tc(x, wp, b, Cin, Cout, dilation=dil, alpha=1.0, n_split=1).double().cpu()
    e_tr = (y1-ref64(x,k,b,dil,trunc_x=True,trunc_w=True)).abs().max().item()
    e_rn = (y1-ref64(x,k,b,dil,round_x=True,trunc_w=True)).abs().max().item()
    print(f"B{B} {H}x{W} Cin{Cin}(cs{cs}) Cout{Cout} d{dil}: err tf32 {out[1]:.2e}  3xtf32 {out[3]:.2e}  direct {ed:.2e} | 1x vs trunc-model {e_tr:.2e} vs rn-model {e_rn:.2e}", flush=True)

case(1,8,16,32,32,1)
case(1,8,16,64,128,1)
case(2,14,32,128,128,1)
case(1,28,64,147,128,1,cs=148)
case(1,28,64,128,96,2)
case(1,24,40,96,64,16)
case(1,7,16,273,128,1,cs=276)
case(2,9,21,64,32,4)
case(1,16,32,192,192,1)

def timeit(B,H,W,Cin,Cout,dil,ns,cs=None,iters=10):
    cs = cs or Cin
    x = torch.randn(B,H,W,cs, device="cuda")[..., :Cin]
    k = torch.randn(3,3,Cin,Cout, device="cuda") / np.sqrt(9*Cin); b = torch.zeros(Cout, device="cuda")
    wp = ops_tc.pack_weights(k); y = torch.empty(B,H,W,Cout, device="cuda")
    f = (lambda: ops_tc.conv3x3_tc(x, wp, b, Cin, Cout, dilation=dil, alpha=0.1, n_split=ns, out=y)) if ns else (lambda: P.ops.conv3x3(x,k,b,dilation=dil,alpha=0.1,out=y))
    for _ in range(3): f()
    s,e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): f()
    e.record(); torch.cuda.synchronize()
    us = s.elapsed_time(e)*1e3/iters
    fl = 2*9*Cin*Cout*B*H*W
    print(f"time B{B} {H}x{W} {Cin}->{Cout} d{dil} split{ns}: {us:.1f} us  {fl/us/1e6:.1f} TFLOP/s", flush=True)
for ns in (0,1,3):
    timeit(8,112,256,147,128,1,ns,cs=148)
    timeit(8,112,256,128,128,1,ns)
    timeit(8,112,256,128,128,4,ns)
    timeit(8,112,256,64,32,1,ns)
    timeit(16,56,128,64,64,1,ns)
