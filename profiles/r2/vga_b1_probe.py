"""Round 2d probe: ONE live frame against the keyframe at 480x640 (kf_vo.py's per-frame call): launch-per-iteration kernels
against the work queue for the finest level(s)."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import make_frame_pairs

dev = torch.device("cuda:0")
C, H, W = 8, 480, 640
data = make_frame_pairs(2, C, H, W, seed=99, n_levels=4, motion=0.05)
key = [{k: lv[k][:1].to(dev).contiguous() for k in ("x0", "s0", "invD0")} for lv in data["levels"]]
lives = [[{k: v[i:i + 1].to(dev).contiguous() for k, v in lv.items() if k in ("x1", "s1", "invD1", "K")} for lv in data["levels"]] for i in range(2)]
pose0 = (data["R0"][:1].to(dev), data["t0"][:1].to(dev))


def solve(i, **kw):
    levels = [dict(kf, **lv) for kf, lv in zip(key, lives[i % 2])]
    return A.uic_solve(levels, pose0, iters=3, remove_tru_sigma=True, shared_keyframe=True, pairwise_extremes=True, **kw)


def lat(**kw):
    for i in range(3):
        solve(i, **kw)
    torch.cuda.synchronize()
    ts = []
    for i in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); solve(i, **kw); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    r = solve(0, timed=True, **kw)
    return round(statistics.median(ts)), [round(sum(r.launch_ms[3 * l:3 * l + 3]) * 1e3) for l in range(4)]


ref = solve(0, queue=False).pose_hist[-1]
print("launch-per-iteration:", lat(queue=False))
for ql in (1, 2):
    for tr in (0, 8, 12, 16, 24):
        kw = dict(queue=True, queue_levels=ql, tile_rows=[0, 0, tr, tr])
        err = (solve(0, **kw).pose_hist[-1] - ref).abs().max().item()
        print(f"queue levels={ql} tile_rows={tr}:", lat(**kw), f"max |pose - lpi| = {err:.1e}")
