"""Round 2e probe: rows per work-queue tile of the 30x40 and 60x80 levels (three levels on the queue, 20 batches per call)."""
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


def run(tr):
    kw = dict(iters=3, remove_tru_sigma=True, group=B, tile_rows=tr)
    for i in range(3):
        A.uic_solve(*sets[i % 2], **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(6):
        A.uic_solve(*sets[i % 2], **kw)
    e1.record()
    torch.cuda.synchronize()
    r = A.uic_solve(*sets[0], timed=True, **kw)
    return e0.elapsed_time(e1) / 6 * 1e3, [round(sum(r.launch_ms[3 * l:3 * l + 3]) * 1e3) for l in range(4)]


for t0 in ((0, 30, 40, 60, 120) if "fine" in sys.argv else ()):
    for t1 in (0, 60):
        t, lv = run([0, 0, t1, t0])
        print(f"G={G} 120x160 tile rows {t0:3d}, 60x80 tile rows {t1:2d}: {t:7.0f} us per call, levels {lv}", flush=True)
if "fine" in sys.argv:
    sys.exit(0)
for t2 in (0, 6, 8, 10, 15, 30):
    t, lv = run([0, t2, 0, 0])
    print(f"G={G} 30x40 tile rows {t2:2d}: {t:7.0f} us per call, levels {lv}", flush=True)
for t1 in (0, 12, 15, 20, 30, 60):
    t, lv = run([0, 0, t1, 0])
    print(f"G={G} 60x80 tile rows {t1:2d}: {t:7.0f} us per call, levels {lv}", flush=True)
