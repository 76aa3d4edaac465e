"""Round 2c probe: 480x640 keyframe tracking, 16 live frames per call, sigma repeated to C channels (check on) and as one map."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import make_frame_pairs

dev = torch.device("cuda:0")
B, C, H, W = 16, 8, 480, 640
for motion in (0.02, 0.05):
    data = make_frame_pairs(B, C, H, W, seed=99, n_levels=4, motion=motion)
    for one_map in (False, True):
        def prep(k, v):
            if k in ("s0", "s1") and not one_map:
                v = v.expand(-1, C, -1, -1)
            return v
        key = [{k: prep(k, lv[k][:1]).to(dev).contiguous() for k in ("x0", "s0", "invD0")} for lv in data["levels"]]
        lives = [[{k: (torch.roll(prep(k, v), s, 0) if k != "K" else v).to(dev).contiguous() for k, v in lv.items() if k in ("x1", "s1", "invD1", "K")}
                  for lv in data["levels"]] for s in range(2)]
        pose0 = (data["R0"].to(dev), data["t0"].to(dev))

        def solve(i, **kw):
            levels = [dict(kf, **lv) for kf, lv in zip(key, lives[i % 2])]
            return A.uic_solve(levels, pose0, iters=3, remove_tru_sigma=True, shared_keyframe=True, pairwise_extremes=True, **kw)

        def timeit(**kw):
            for i in range(2):
                solve(i, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(6):
                solve(i, **kw)
            e1.record()
            torch.cuda.synchronize()
            r = solve(0, timed=True, **kw)
            return round(e0.elapsed_time(e1) / 6 * 1e3), [round(sum(r.launch_ms[3 * l:3 * l + 3]) * 1e3) for l in range(4)]

        print(f"motion {motion} {'one map passed in' if one_map else 'sigma repeated    '}: queue {timeit(queue=True)}  launch-per-iteration {timeit(queue=False)}", flush=True)
