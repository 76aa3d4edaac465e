"""ncu target: two training steps (solver forward + backward) of the bench workload."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import levels_to, make_frame_pairs

data = make_frame_pairs(64, 8, 120, 160, seed=1234, n_levels=4)
lv = levels_to(data["levels"], "cuda:0")
for d in lv:
    for k in ("x0", "x1", "s0", "s1"):
        d[k].requires_grad_(True)
R = data["R0"].cuda().requires_grad_(True)
t = data["t0"].cuda().requires_grad_(True)
for _ in range(2):
    outs = A.uic_track(lv, (R, t), iters=3, remove_tru_sigma=True, check=False)
    sum(Rl.sum() + tl.sum() for Rl, tl, _ in outs).backward()
torch.cuda.synchronize()
print("ok")
