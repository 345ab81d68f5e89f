"""bench.py's level-2 cost-volume roofline measurement alone (ncu target / quick check): python tools/roofline_once.py [B]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pwcnet_b200 as P
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.cuda.set_device(0)
for reps in (12, 50):
    r = bench.roofline_cost_volume(P, torch, torch.device("cuda:0"), B, reps=reps)
    print(json.dumps({k: r[k] for k in ("us_per_launch", "achieved", "frac", "launches_timed")}))
