"""Round 2 probe: 480x640 keyframe tracking, 16 live frames per call: work-queue vs launch-per-iteration, tile heights, motion."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import make_frame_pairs

dev = torch.device("cuda:0")
B, C, H, W = 16, 8, 480, 640
for motion in (0.02, 0.05):
    data = make_frame_pairs(B, C, H, W, seed=99, n_levels=4, motion=motion)
    key = [{k: lv[k][:1].to(dev).contiguous() for k in ("x0", "s0", "invD0")} for lv in data["levels"]]
    lives = [[{k: (torch.roll(v, s, 0) if k != "K" else v).to(dev).contiguous() for k, v in lv.items() if k in ("x1", "s1", "invD1", "K")}
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
        return e0.elapsed_time(e1) / 6 * 1e3, [round(sum(r.launch_ms[3 * l:3 * l + 3]) * 1e3) for l in range(4)], r.queue_kernel_ms

    print(f"motion {motion}", flush=True)
    print("  launch-per-iteration:", timeit(queue=False), flush=True)
    for tr in (0, 16, 24, 32, 48):
        print(f"  queue tile_rows={tr}:", timeit(queue=True, tile_rows=[0, 0, 0, tr]), flush=True)
    print("  queue, 2 finest levels:", timeit(queue=True, queue_levels=2), flush=True)
    print("  queue, generic tiling knobs: ctas=296:", timeit(queue=True, queue_ctas=296), flush=True)
