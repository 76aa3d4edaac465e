"""ncu target: keyframe tracking at 480x640, 16 live frames against one keyframe (per-frame sigma extremes), sigma
repeated to C channels as the reference does: a few calls, the finest level as a work-queue launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import make_frame_pairs

dev = torch.device("cuda:0")
B, C, H, W = 16, 8, 480, 640
data = make_frame_pairs(B, C, H, W, seed=99, n_levels=4, motion=0.05)
key = [{k: lv[k][:1].to(dev).contiguous() for k in ("x0", "s0", "invD0")} for lv in data["levels"]]
live = [{k: v.to(dev).contiguous() for k, v in lv.items() if k in ("x1", "s1", "invD1", "K")} for lv in data["levels"]]
tracker = A.KeyframeTracker(key, iters=3, remove_tru_sigma=True)
pose0 = (data["R0"].to(dev), data["t0"].to(dev))
for i in range(4):
    r = tracker.track(live, pose0)
torch.cuda.synchronize()
r.raise_if_bad()
print("ok")
