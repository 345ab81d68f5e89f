"""use_dc=True training: per-tensor gradient error of the 3xf16 path vs the oracle (debug aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pwcnet_b200 as P
from pwcnet_b200.train import Trainer
from oracle import pwc_oracle as O
W = O.glorot_weights(5, gain=1.0, bias_scale=0.02, use_dc=True)
im0, im1 = O.synthetic_pair(2, 64, 128, 4, shift=(-4, 2))
gt = np.random.default_rng(1).normal(0, 5, (2, 64, 128, 2)).astype(np.float32)
Wt = {k: torch.from_numpy(v.copy()).requires_grad_(True) for k, v in W.items()}
total, epe, ff, pyr = O.training_loss(Wt, im0, im1, gt, gamma=0.0, use_dc=True)
total.backward()
for kw in (dict(tc_dgrad=True, tc_wgrad=True), dict(tc_dgrad=False, tc_wgrad=True), dict(tc_dgrad=True, tc_wgrad=False)):
    model = P.PWCDCNet(weights=W, use_dc=True, precision="3xf16")
    tr = Trainer(model, **kw)
    tr.forward_backward(im0, im1, gt)
    torch.cuda.synchronize()
    errs = []
    for name in model.var_names:
        got, r = tr.grads[name].cpu().numpy(), Wt[name].grad.numpy()
        errs.append((float(np.abs(got - r).max()) / float(np.abs(r).max()), name, tuple(r.shape)))
    errs.sort(reverse=True)
    print(kw, [f"{e:.1e} {n.split('pwcdcnet/')[1]} {s}" for e, n, s in errs[:6]])
