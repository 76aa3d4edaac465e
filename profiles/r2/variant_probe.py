"""One library variant (DPFT_LIB_PATH): finest-level work-queue launch at G batches of 64 pairs, 120x160."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import make_frame_pairs

dev = torch.device("cuda:0")
B, C, H, W = 64, 8, 120, 160
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
parts = [make_frame_pairs(B, C, H, W, seed=1234 + g, n_levels=4) for g in range(G)]
levels = [{k: torch.cat([p["levels"][l][k] for p in parts]).to(dev) for k in parts[0]["levels"][l]} for l in range(4)]
pose = (torch.cat([p["R0"] for p in parts]).to(dev), torch.cat([p["t0"] for p in parts]).to(dev))
sets = [([{k: torch.roll(v, s * 5, 0).contiguous() for k, v in lv.items()} for lv in levels], pose) for s in range(2)]
out = []
for tr in (30, 40):
    kw = dict(iters=3, remove_tru_sigma=True, group=B, queue=True, tile_rows=[0, 0, 0, tr])
    for i in range(3):
        A.uic_solve(*sets[i % 2], **kw)
    torch.cuda.synchronize()
    n = 6
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        A.uic_solve(*sets[i % 2], **kw)
    e1.record()
    torch.cuda.synchronize()
    lv0 = []
    for i in range(4):
        r = A.uic_solve(*sets[i % 2], timed=True, **kw)
        lv0.append(sum(r.launch_ms[9:12]) * 1e3)
    out.append(f"tr={tr}: {e0.elapsed_time(e1) * 1e3 / n / G:6.1f} us/batch, level-0 {min(lv0) / G / 3:5.2f} us per batch-iteration")
print(os.path.basename(os.environ.get("DPFT_LIB_PATH", "default")), " | ".join(out), flush=True)
