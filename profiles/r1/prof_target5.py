"""ncu target: one training step (forward + backward through 4 levels x 3 iterations)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import levels_to, make_frame_pairs

data = make_frame_pairs(64, 8, 120, 160, seed=1234, n_levels=4)
lv = levels_to(data["levels"], "cuda:0")
for l in lv:
    for k in ("x0", "x1", "s0", "s1"):
        l[k].requires_grad_(True)
pose = (data["R0"].cuda(), data["t0"].cuda())
for _ in range(2):
    outs = A.uic_track(lv, pose, iters=3, remove_tru_sigma=True)
    loss = sum((R.sum() + t.sum()) for R, t, _ in outs)
    loss.backward()
torch.cuda.synchronize()
print("ok")
