"""ncu target: two calls of G batches (64 pairs, 120x160, sigma repeated to C channels, replication found on the device)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import make_frame_pairs

dev = torch.device("cuda:0")
B, C, H, W = 64, 8, 120, 160
G = int(sys.argv[1]) if len(sys.argv) > 1 else 20
parts = [make_frame_pairs(B, C, H, W, seed=1234 + g, n_levels=4) for g in range(G)]
levels = [{k: torch.cat([p["levels"][l][k] for p in parts]).to(dev) for k in parts[0]["levels"][l]} for l in range(4)]
for lv in levels:
    lv["s0"] = lv["s0"].expand(-1, C, -1, -1).contiguous(); lv["s1"] = lv["s1"].expand(-1, C, -1, -1).contiguous()
pose = (torch.cat([p["R0"] for p in parts]).to(dev), torch.cat([p["t0"] for p in parts]).to(dev))
for i in range(2):
    r = A.uic_solve(levels, pose, iters=3, remove_tru_sigma=True, group=B)
torch.cuda.synchronize()
r.raise_if_bad()
print("ok")
