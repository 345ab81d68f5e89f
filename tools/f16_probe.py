"""fp16x3 conv: accuracy vs fp64 and timing vs 3xtf32."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pwcnet_b200 as P
from pwcnet_b200 import ops_tc
torch.manual_seed(0)
def ref64(x,k,b,dil,stride):
    x=x.double().cpu(); k=k.double().cpu()
    H,W=x.shape[1],x.shape[2]
    def pad(n):
        o=-(-n//stride); t=max((o-1)*stride+2*dil+1-n,0); return t//2, t-t//2
    (pt,pb),(pl,pr)=pad(H),pad(W)
    xn=torch.nn.functional.pad(x.permute(0,3,1,2),(pl,pr,pt,pb))
    y=torch.nn.functional.conv2d(xn,k.permute(3,2,0,1),b.double().cpu(),stride=stride,dilation=dil)
    return y.permute(0,2,3,1)
def case(B,H,W,Cin,Cout,dil,cs=None,stride=1,scale=1.0):
    cs=cs or Cin
    x=(torch.randn(B,H,W,cs,device="cuda")*scale)[...,:Cin]
    k=torch.randn(3,3,Cin,Cout,device="cuda")/np.sqrt(9*Cin); b=torch.randn(Cout,device="cuda")*0.1
    r=ref64(x,k,b,dil,stride); r=torch.maximum(0.1*r,r)
    y=ops_tc.conv3x3_tc_f16(x,ops_tc.pack_weights_f16(k),b,Cin,Cout,dilation=dil,alpha=0.1,stride=stride)
    torch.cuda.synchronize()
    e16=(y.double().cpu()-r).abs().max().item()
    e32=float('nan')
    if Cin==16 or Cin>=32:
        y2=ops_tc.conv3x3_tc(x,ops_tc.pack_weights(k),b,Cin,Cout,dilation=dil,alpha=0.1,n_split=3,stride=stride)
        e32=(y2.double().cpu()-r).abs().max().item()
    yd=P.ops.conv3x3(x,k,b,stride=stride,dilation=dil,alpha=0.1)
    ed=(yd.double().cpu()-r).abs().max().item()
    print(f"B{B} {H}x{W} Cin{Cin}(cs{cs}) Cout{Cout} d{dil} s{stride} scale{scale}: err 3xf16 {e16:.2e}  3xtf32 {e32:.2e}  direct {ed:.2e}  max|y| {r.abs().max().item():.1f}",flush=True)
case(1,8,16,32,32,1)
case(2,14,32,128,128,1)
case(1,28,64,147,128,1,cs=148)
case(1,7,16,273,128,1,cs=276)
case(1,28,64,128,96,2)
case(1,24,40,96,64,16)
case(2,9,21,64,32,4)
case(1,16,32,192,192,1)
case(2,32,48,16,16,1)
case(2,32,48,16,32,1,stride=2)
case(1,15,17,32,64,1,stride=2)
case(1,28,64,34,128,1,cs=36)
case(2,14,32,128,128,1,scale=100.0)
case(2,14,32,128,128,1,scale=1e-3)
def timeit(B,H,W,Cin,Cout,dil,mode,cs=None,iters=10):
    cs=cs or Cin
    x=torch.randn(B,H,W,cs,device="cuda")[...,:Cin]
    k=torch.randn(3,3,Cin,Cout,device="cuda")/np.sqrt(9*Cin); b=torch.zeros(Cout,device="cuda"); y=torch.empty(B,H,W,Cout,device="cuda")
    if mode=="f16":
        wp=ops_tc.pack_weights_f16(k); f=lambda: ops_tc.conv3x3_tc_f16(x,wp,b,Cin,Cout,dilation=dil,alpha=0.1,out=y)
    else:
        wp=ops_tc.pack_weights(k); f=lambda: ops_tc.conv3x3_tc(x,wp,b,Cin,Cout,dilation=dil,alpha=0.1,n_split=3,out=y)
    for _ in range(3): f()
    s,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): f()
    e.record(); torch.cuda.synchronize()
    us=s.elapsed_time(e)*1e3/iters
    print(f"time B{B} {H}x{W} {Cin}->{Cout} d{dil} {mode}: {us:.1f} us  {2*9*Cin*Cout*B*H*W/us/1e6:.1f} TFLOP/s",flush=True)
for mode in ("tf32x3","f16"):
    timeit(8,112,256,147,128,1,mode,cs=148)
    timeit(8,112,256,128,128,1,mode)
    timeit(8,112,256,128,128,4,mode)
    timeit(8,112,256,64,32,1,mode)
    timeit(16,56,128,64,64,1,mode)
    timeit(16,112,256,32,32,1,mode)
    timeit(16,224,512,16,16,1,mode)
