"""ncu target: two solves with the staged-footprint kernel, one launch per iteration."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import levels_to, make_frame_pairs

data = make_frame_pairs(64, 8, 120, 160, seed=1234, n_levels=4)
lv = levels_to(data["levels"], "cuda:0")
pose = (data["R0"].cuda(), data["t0"].cuda())
for _ in range(2):
    res = A.uic_solve(lv, pose, iters=3, remove_tru_sigma=True, staged_footprint=True)
torch.cuda.synchronize()
res.raise_if_bad()
print("ok")
