"""Level-0 launch time against the batch size (tile height follows from the wave-aware plan): separates the fixed
cost of a launch (prologue, reduction tail, solves) from the per-row cost of the tile walk."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import levels_to, make_frame_pairs

staged = "--staged" in sys.argv
for B in (4, 8, 16, 32, 64, 128):
    data = make_frame_pairs(B, 8, 120, 160, seed=1234, n_levels=4)
    lv = levels_to(data["levels"], "cuda:0")
    pose = (data["R0"].cuda(), data["t0"].cuda())
    best = None
    for _ in range(6):
        res = A.uic_solve(lv, pose, iters=3, remove_tru_sigma=True, staged_footprint=staged, timed=True)
        torch.cuda.synchronize()
        ms = res.launch_ms
        best = ms if best is None else [min(a, b) for a, b in zip(best, ms)]
    print(B, "staged" if staged else "plain", [round(x * 1e3, 1) for x in best])
