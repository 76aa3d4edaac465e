"""Round 2e probe: 480x640 keyframe tracking, 16 live frames per call: rows per work-queue tile of the finest level."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import make_frame_pairs

dev = torch.device("cuda:0")
C, H, W = 8, 480, 640
for B in (16, 64):
    data = make_frame_pairs(16, C, H, W, seed=99, n_levels=4)
    rep = B // 16
    prep = lambda k, v: v.expand(-1, C, -1, -1) if k in ("s0", "s1") else v
    key = [{k: prep(k, lv[k][:1]).to(dev).contiguous() for k in ("x0", "s0", "invD0")} for lv in data["levels"]]
    lives = [[{k: (torch.roll(prep(k, v), s, 0) if k != "K" else v).repeat(rep, 1, 1, 1).to(dev).contiguous() if v.dim() == 4 else v.repeat(rep, 1).to(dev)
               for k, v in lv.items() if k in ("x1", "s1", "invD1", "K")} for lv in data["levels"]] for s in range(2)]
    pose0 = (data["R0"].repeat(rep, 1, 1).to(dev), data["t0"].repeat(rep, 1).to(dev))

    def solve(i, **kw):
        levels = [dict(kf, **lv) for kf, lv in zip(key, lives[i % 2])]
        return A.uic_solve(levels, pose0, iters=3, remove_tru_sigma=True, shared_keyframe=True, pairwise_extremes=True, queue=True, **kw)

    for tr in [int(x) for x in sys.argv[1:]] or (0, 8, 12, 16, 20, 24, 30, 32, 40, 48):
        kw = dict(tile_rows=[0, 0, 0, tr])
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
        lv = [round(sum(r.launch_ms[3 * l:3 * l + 3]) * 1e3) for l in range(4)]
        print(f"B={B:3d} finest-level tile rows {tr:2d}: {e0.elapsed_time(e1) / 6 * 1e3:8.0f} us per call, levels {lv}, queue kernel {r.queue_kernel_ms[-1] * 1e3:.0f} us", flush=True)
