"""Small fixed program for ncu captures: two coarse-to-fine U_IC solves of the bench workload
(B=64, C=8, 120x160, 4 levels x 3 iterations).  Usage: python profiles/prof_target.py [B H W]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import levels_to, make_frame_pairs

B, H, W = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (64, 120, 160)
data = make_frame_pairs(B, 8, H, W, seed=1234, n_levels=4)
lv = levels_to(data["levels"], "cuda:0")
pose = (data["R0"].cuda(), data["t0"].cuda())
for _ in range(2):
    res = A.uic_solve(lv, pose, iters=3, remove_tru_sigma=True)
torch.cuda.synchronize()
res.raise_if_bad()
print("ok", res.pose[1][0].tolist())
