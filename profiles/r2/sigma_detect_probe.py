"""Round 2 probe: a call of G batches (64 pairs, 120x160, C = 8, sigma repeated to C channels as the reference does) with the
device-side replication check on (default), off, and with the caller passing the one map itself."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import make_frame_pairs

dev = torch.device("cuda:0")
B, C, H, W = 64, 8, 120, 160
G = int(sys.argv[1]) if len(sys.argv) > 1 else 20
sets = []
for s in range(2):
    parts = [make_frame_pairs(B, C, H, W, seed=1234 + 17 * s + g, n_levels=4) for g in range(G)]
    levels = [{k: torch.cat([p["levels"][l][k] for p in parts]).to(dev) for k in parts[0]["levels"][l]} for l in range(4)]
    for lv in levels:
        lv["s0"] = lv["s0"].expand(-1, C, -1, -1).contiguous(); lv["s1"] = lv["s1"].expand(-1, C, -1, -1).contiguous()
    pose = (torch.cat([p["R0"] for p in parts]).to(dev), torch.cat([p["t0"] for p in parts]).to(dev))
    sets.append((levels, pose))
one = [([dict(lv, s0=lv["s0"][:, :1].contiguous(), s1=lv["s1"][:, :1].contiguous()) for lv in s[0]], s[1]) for s in sets]


def timeit(fn, n=6, warm=3):
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


kw = dict(iters=3, remove_tru_sigma=True, group=B)
if len(sys.argv) > 2:
    kw["queue_levels"] = int(sys.argv[2])          # how many of the finest levels run as work-queue launches
for name, data, tun in (("check on (default)", sets, None), ("check off          ", sets, dict(sigma_detect=1)), ("one map passed in  ", one, None),
                        ("check off, generic geometry", sets, dict(sigma_detect=1, generic_geometry=1))):
    t = timeit(lambda i: A.uic_solve(*data[i % 2], tuning=tun, **kw))
    r = A.uic_solve(*data[0], timed=True, tuning=tun, **kw)
    its = [round(x * 1e3, 1) for x in r.launch_ms]
    print(f"G={G} {name}: {t / G:7.1f} us per batch ({t:8.1f} us per call, {B * G / t * 1e3:8.1f} k pairs/s); level sums us "
          f"{[round(sum(its[3 * l:3 * l + 3]), 1) for l in range(4)]}", flush=True)
