"""One library build (DPFT_LIB_PATH or the default): default-policy calls of G batches of 64 pairs (120x160 pyramid, sigma
repeated to C channels) and the 480x640 keyframe call of 16 frames -- us per call and per level.
Usage: [DPFT_LIB_PATH=profiles/r2/variants/X.so] python profiles/r2/ab_probe.py [batches ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import make_frame_pairs

dev = torch.device("cuda:0")
name = os.path.basename(os.environ.get("DPFT_LIB_PATH", "default"))
B, C, H, W = 64, 8, 120, 160
groups = [int(x) for x in sys.argv[1:]] or [8, 20]
gmax = max(groups)
sets = []
for s in range(2):
    parts = [make_frame_pairs(B, C, H, W, seed=1234 + 17 * s + g, n_levels=4) for g in range(gmax)]
    levels = [{k: torch.cat([p["levels"][l][k] for p in parts]).to(dev) for k in parts[0]["levels"][l]} for l in range(4)]
    for lv in levels:
        lv["s0"] = lv["s0"].expand(-1, C, -1, -1).contiguous(); lv["s1"] = lv["s1"].expand(-1, C, -1, -1).contiguous()
    pose = (torch.cat([p["R0"] for p in parts]).to(dev), torch.cat([p["t0"] for p in parts]).to(dev))
    sets.append((levels, pose))


def measure(fn, n):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    best = 1e30
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return best


for G in groups:
    data = [([{k: v[:B * G] for k, v in lv.items()} for lv in s[0]], (s[1][0][:B * G], s[1][1][:B * G])) for s in sets]
    kw = dict(iters=3, remove_tru_sigma=True, group=B)
    t = measure(lambda i: A.uic_solve(*data[i % 2], **kw), max(4, 60 // G))
    r = A.uic_solve(*data[0], timed=True, **kw)
    lv = [round(sum(r.launch_ms[3 * l:3 * l + 3]) * 1e3) for l in range(4)]
    print(f"{name}: 120x160 G={G:2d}: {t:7.0f} us per call, levels {lv}, queue kernels {[round(x * 1e3) for x in r.queue_kernel_ms]}", flush=True)
del sets
torch.cuda.empty_cache()
data = make_frame_pairs(16, C, 480, 640, seed=99, n_levels=4)
prep = lambda k, v: v.expand(-1, C, -1, -1) if k in ("s0", "s1") else v
key = [{k: prep(k, lv[k][:1]).to(dev).contiguous() for k in ("x0", "s0", "invD0")} for lv in data["levels"]]
lives = [[{k: (torch.roll(prep(k, v), s, 0) if k != "K" else v).to(dev).contiguous() for k, v in lv.items() if k in ("x1", "s1", "invD1", "K")}
          for lv in data["levels"]] for s in range(2)]
pose0 = (data["R0"].to(dev), data["t0"].to(dev))


def vga(i, **kw):
    levels = [dict(kf, **lv) for kf, lv in zip(key, lives[i % 2])]
    return A.uic_solve(levels, pose0, iters=3, remove_tru_sigma=True, shared_keyframe=True, pairwise_extremes=True, queue=True, **kw)


t = measure(vga, 6)
r = vga(0, timed=True)
print(f"{name}: 480x640 16 frames: {t:7.0f} us per call, queue kernel {r.queue_kernel_ms[-1] * 1e3:.0f} us", flush=True)
