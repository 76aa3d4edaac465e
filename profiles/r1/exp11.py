"""In-kernel timeline of ONE coarse-level launch (needs -DDPFT_DEBUG_STAMPS): where do the ~20 us of a launch that
touches 300 pixels per pair go?"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deep_prob_feature_track_b200 import _lib, algorithms as A
from deep_prob_feature_track_b200.synthetic import levels_to, make_frame_pairs

data = make_frame_pairs(64, 8, 120, 160, seed=1234, n_levels=4)
lv = levels_to(data["levels"], "cuda:0")
pose = (data["R0"].cuda(), data["t0"].cuda())
names = ["resident", "prev done", "tile walked", "cta reduced", "pair last cta", "pair reduced", "extremes", "solved"]
for nl in (1, 2):
    for rep in range(3):
        res = A.uic_solve(lv[:nl], pose, iters=3, remove_tru_sigma=True, timed=True)
        torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * 16)()
        _lib.lib().dpft_debug_read_stamps(buf)
        t0 = buf[0]
        print(f"level {lv[nl - 1]['x0'].shape[2]}x{lv[nl - 1]['x0'].shape[3]} launch_us", round(res.launch_ms[-1] * 1e3, 1),
              " ".join(f"{n}={((buf[i] - t0) / 1e3):.1f}" for i, n in enumerate(names)))
