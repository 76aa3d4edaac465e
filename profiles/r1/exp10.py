"""Level launch times against the channel count (C = 1 is the reference's default feature_channel)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import levels_to, make_frame_pairs

for C in (1, 2, 4, 8):
    data = make_frame_pairs(64, C, 120, 160, seed=1234, n_levels=4)
    lv = levels_to(data["levels"], "cuda:0")
    pose = (data["R0"].cuda(), data["t0"].cuda())
    best = None
    for _ in range(6):
        res = A.uic_solve(lv, pose, iters=3, remove_tru_sigma=True, timed=True)
        torch.cuda.synchronize()
        ms = res.launch_ms
        best = ms if best is None else [min(a, b) for a, b in zip(best, ms)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        A.uic_solve(lv, pose, iters=3, remove_tru_sigma=True)
    e1.record()
    torch.cuda.synchronize()
    lvl0 = sum(best[-3:]) / 3
    print(f"C={C} step {e0.elapsed_time(e1) / 50:.3f} ms  level-0 launch {lvl0 * 1e3:.1f} us = "
          f"{(4 * C + 2) * 4 * 19200 * 64 / (lvl0 * 1e-3) / 1e9:.0f} GB/s algorithmic", [round(x * 1e3, 1) for x in best])
