"""Round 2 probe: launch-per-iteration path vs the work-queue launch, 1..8 batches of 64 pairs per launch.
Usage: python profiles/r2/queue_probe.py [groups ...]   (prints one line per configuration)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import make_frame_pairs

dev = torch.device("cuda:0")
B, C, H, W = 64, 8, 120, 160
groups = [int(x) for x in sys.argv[1:]] or [1, 2, 4]
gmax = max(groups)
sets = []
for s in range(2):
    parts = [make_frame_pairs(B, C, H, W, seed=1234 + 17 * s + g, n_levels=4) for g in range(gmax)]
    levels = [{k: torch.cat([p["levels"][l][k] for p in parts]).to(dev) for k in parts[0]["levels"][l]} for l in range(4)]
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


lv1 = [sub(*s, B) for s in sets]
t = timeit(lambda i: A.uic_solve(*lv1[i % 2], iters=3, remove_tru_sigma=True, queue=False))
print(f"launch-per-iteration, 1 stream: {t:8.1f} us per batch of {B}", flush=True)
r = A.uic_solve(*lv1[0], iters=3, remove_tru_sigma=True, queue=False, timed=True)
print("   per-iteration launch us:", [round(x * 1e3, 1) for x in r.launch_ms], flush=True)
for G in groups:
    data = [sub(*s, B * G) for s in sets]
    for tr in (None, [0, 0, 0, 12], [0, 0, 0, 20], [0, 0, 0, 40]):
        kw = dict(iters=3, remove_tru_sigma=True, group=B, tile_rows=tr)
        t = timeit(lambda i: A.uic_solve(*data[i % 2], **kw), n=max(4, 20 // G))
        r = A.uic_solve(*data[0], timed=True, **kw)
        its = [round(x * 1e3, 1) for x in r.launch_ms]
        print(f"queue G={G} tile_rows={tr}: {t / G:8.1f} us per batch ({t:8.1f} us per launch); level sums us "
              f"{[round(sum(its[3 * l:3 * l + 3]), 1) for l in range(4)]}", flush=True)
    ref = A.uic_solve(*lv1[0], iters=3, remove_tru_sigma=True, queue=False)
    q = A.uic_solve(*data[0], iters=3, remove_tru_sigma=True, group=B)
    torch.cuda.synchronize()
    print(f"   max |pose diff| vs launch-per-iteration (first batch): {(q.pose_hist[:, :B] - ref.pose_hist).abs().max().item():.2e}", flush=True)
print("--- queue_levels / larger calls", flush=True)
for G in groups:
    if G < 4:
        continue
    data = [sub(*s, B * G) for s in sets]
    for ql in (1, 2, 3):
        kw = dict(iters=3, remove_tru_sigma=True, group=B, queue=True, queue_levels=ql)
        t = timeit(lambda i: A.uic_solve(*data[i % 2], **kw), n=max(4, 20 // G))
        r = A.uic_solve(*data[0], timed=True, **kw)
        its = [round(x * 1e3, 1) for x in r.launch_ms]
        print(f"queue G={G} queue_levels={ql}: {t / G:8.1f} us per batch; level sums us {[round(sum(its[3 * l:3 * l + 3]), 1) for l in range(4)]}", flush=True)
