"""ncu target: one batch of 64 pairs (120x160, C = 8) through uic_solve with the point-to-plane ICP term (config 5)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import make_frame_pairs

dev = torch.device("cuda:0")
d = make_frame_pairs(64, 8, 120, 160, seed=5, n_levels=4, with_depth=True)
levels = [{k: v.to(dev) for k, v in lv.items()} for lv in d["levels"]]
pose = (d["R0"].to(dev), d["t0"].to(dev))
for i in range(3):
    r = A.uic_solve(levels, pose, iters=3, remove_tru_sigma=True, combine_icp=True, w_icp=0.01)
torch.cuda.synchronize()
r.raise_if_bad()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10):
    A.uic_solve(levels, pose, iters=3, remove_tru_sigma=True, combine_icp=True, w_icp=0.01)
e1.record(); torch.cuda.synchronize()
print(f"with ICP term: {e0.elapsed_time(e1) / 10 * 1e3:.0f} us per solve of 64 pairs")
e0.record()
for i in range(10):
    A.uic_solve(levels, pose, iters=3, remove_tru_sigma=True)
e1.record(); torch.cuda.synchronize()
print(f"without:       {e0.elapsed_time(e1) / 10 * 1e3:.0f} us per solve of 64 pairs")
