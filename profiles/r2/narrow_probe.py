"""Round 2e probe: how many pyramid levels run as work-queue launches, with the staged routine's NARROW form
(30x40 / 15x20 levels) against the plain tile routine on the same queue, and against the default policy.
Usage: python profiles/r2/narrow_probe.py [batches ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import make_frame_pairs

dev = torch.device("cuda:0")
B, C, H, W = 64, 8, 120, 160
groups = [int(x) for x in sys.argv[1:]] or [8, 20]
gmax = max(groups)
sets = []
for s in range(2):
    parts = [make_frame_pairs(B, C, H, W, seed=1234 + 17 * s + g, n_levels=4) for g in range(gmax)]
    levels = [{k: torch.cat([p["levels"][l][k] for p in parts]).to(dev) for k in parts[0]["levels"][l]} for l in range(4)]
    for lv in levels:     # the bench's format: sigma repeated to C channels
        lv["s0"] = lv["s0"].expand(-1, C, -1, -1).contiguous(); lv["s1"] = lv["s1"].expand(-1, C, -1, -1).contiguous()
    pose = (torch.cat([p["R0"] for p in parts]).to(dev), torch.cat([p["t0"] for p in parts]).to(dev))
    sets.append((levels, pose))


def sub(levels, pose, n):
    return [{k: v[:n] for k, v in lv.items()} for lv in levels], (pose[0][:n], pose[1][:n])


def timeit(fn, n=20, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3   # us


variants = [("default policy      ", dict()),
            ("queue_levels=1       ", dict(queue_levels=1)),
            ("queue_levels=2       ", dict(queue_levels=2)),
            ("queue_levels=3 narrow", dict(queue_levels=3)),
            ("queue_levels=3 plain ", dict(queue_levels=3, tuning=dict(small_levels=2))),
            ("queue_levels=4 narrow", dict(queue_levels=4)),
            ("queue_levels=4 plain ", dict(queue_levels=4, tuning=dict(small_levels=2)))]
for G in groups:
    data = [sub(*s, B * G) for s in sets]
    ref = None
    for name, extra in variants:
        kw = dict(iters=3, remove_tru_sigma=True, group=B, **extra)
        try:
            t = timeit(lambda i: A.uic_solve(*data[i % 2], **kw), n=max(4, 60 // G))
            r = A.uic_solve(*data[0], timed=True, **kw)
            its = [round(x * 1e3, 1) for x in r.launch_ms]
            r = A.uic_solve(*data[0], **kw)
            torch.cuda.synchronize()
            if ref is None:
                ref = r
            dp = (r.pose_hist - ref.pose_hist).abs().max().item()
            ds = ((r.sys_hist - ref.sys_hist).abs().amax(dim=(1, 2)) / ref.sys_hist.abs().amax(dim=(1, 2))).max().item()
            print(f"G={G:2d} {name}: {t / G:8.1f} us per batch ({t:8.1f} us per call); level sums us "
                  f"{[round(sum(its[3 * l:3 * l + 3]), 1) for l in range(4)]}; vs default: |pose| {dp:.2e} sys rel {ds:.2e} "
                  f"status {int(r.status.item())}", flush=True)
        except Exception as e:   # noqa: BLE001
            print(f"G={G:2d} {name}: FAILED {type(e).__name__}: {e}", flush=True)
