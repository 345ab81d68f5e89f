"""2-rank NCCL check of the bucketed gradient all-reduce with the weight gradients on a side stream: the all-reduced gradient
(sum over ranks / world) of a batch split over ranks must equal the whole-batch gradient computed by one rank.
torchrun --nproc-per-node 2 tools/ddp_grad_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
import pwcnet_b200 as P
from oracle import pwc_oracle as O
W = O.glorot_weights(7, gain=1.4, bias_scale=0.02)
im0, im1, flow = O.synthetic_textured_pair(2 * world, 128, 256, 9, 6.0)
gt = (flow + np.random.default_rng(1).normal(0, 2, flow.shape)).astype(np.float32)
# every rank: its slice, with the process group
tr = P.Trainer(P.PWCDCNet(weights=W))             # default process group: the gradient buckets are all-reduced over both ranks
sl = slice(2 * rank, 2 * rank + 2)
for _ in range(3):                                  # repeated: the side stream / bucket events of consecutive steps
    tr.forward_backward(im0[sl], im1[sl], gt[sl])
    w = tr._allreduce_finish()
    torch.cuda.synchronize()
g_dist = (tr.grad_flat / w).clone()
if rank == 0:
    t1 = P.Trainer(P.PWCDCNet(weights=W))
    t1.skip_allreduce = True                       # one rank, whole batch
    t1.forward_backward(im0, im1, gt)
    torch.cuda.synchronize()
    ref = t1.grad_flat
    err = float((g_dist - ref).abs().max()) / float(ref.abs().max())
    print(f"world {w}: max-abs gradient difference / max|grad| = {err:.3e}")
    assert w == world and err < 1e-4, err
    print("DDP GRAD CHECK OK")
dist.barrier()
dist.destroy_process_group()
