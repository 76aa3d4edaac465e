"""ncu target: a few finest-level work-queue launches (G batches of 64 pairs, 120x160, C = 8, remove_tru_sigma)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import make_frame_pairs

dev = torch.device("cuda:0")
B, C, H, W = 64, 8, 120, 160
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
parts = [make_frame_pairs(B, C, H, W, seed=1234 + g, n_levels=4) for g in range(G)]
levels = [{k: torch.cat([p["levels"][l][k] for p in parts]).to(dev) for k in parts[0]["levels"][l]} for l in range(4)]
pose = (torch.cat([p["R0"] for p in parts]).to(dev), torch.cat([p["t0"] for p in parts]).to(dev))
for i in range(4):
    r = A.uic_solve(levels, pose, iters=3, remove_tru_sigma=True, group=B, queue=True, tile_rows=[0, 0, 0, 40])
torch.cuda.synchronize()
r.raise_if_bad()
print("ok")
