"""In-kernel timeline of the last (finest-level) launch of a solve; needs a library built with -DDPFT_DEBUG_STAMPS
(DPFT_NVCC_EXTRA=-DDPFT_DEBUG_STAMPS).  Prints microseconds relative to 'CTA resident'."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deep_prob_feature_track_b200 import _lib, algorithms as A
from deep_prob_feature_track_b200.synthetic import levels_to, make_frame_pairs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
data = make_frame_pairs(B, 8, 120, 160, seed=1234, n_levels=4)
lv = levels_to(data["levels"], "cuda:0")
pose = (data["R0"].cuda(), data["t0"].cuda())
names = ["resident", "prev done", "tile walked", "cta reduced", "pair last cta", "pair reduced", "extremes", "solved"]
for rep in range(4):
    res = A.uic_solve(lv, pose, iters=3, remove_tru_sigma=True, timed=True)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 16)()
    _lib.lib().dpft_debug_read_stamps(buf)
    t0 = buf[0]
    print(B, "launch_ms", round(res.launch_ms[-1] * 1e3, 1), " ".join(f"{n}={((buf[i] - t0) / 1e3):.1f}" for i, n in enumerate(names)))
