import os, sys
sys.path.insert(0, "/root/repo")
import torch, statistics
from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import make_frame_pairs
dev = torch.device("cuda:0")
B, C, H, W, G = 64, 8, 120, 160, 8
sets = []
for s in range(2):
    parts = [make_frame_pairs(B, C, H, W, seed=1234 + 17 * s + g, n_levels=4) for g in range(G)]
    levels = [{k: torch.cat([p["levels"][l][k] for p in parts]).to(dev) for k in parts[0]["levels"][l]} for l in range(4)]
    pose = (torch.cat([p["R0"] for p in parts]).to(dev), torch.cat([p["t0"] for p in parts]).to(dev))
    sets.append((levels, pose))
full = A.uic_solve(*sets[0], iters=3, remove_tru_sigma=True, group=B, queue=True)
pose_l0 = A.unpack_pose(full.pose_hist[9])
fine = [s[0][-1] for s in sets]
for tr in (0, 24, 30, 40, 60, 120):
    for qc in (0, 296, 430):
        ms = []
        for i in range(8):
            r = A.uic_solve([fine[i % 2]], pose_l0, iters=3, remove_tru_sigma=True, group=B, queue=True, timed=True, tile_rows=[tr], queue_ctas=qc)
            if i >= 2: ms.append(r.queue_kernel_ms[0])
        print(f"tile_rows={tr:3d} queue_ctas={qc:3d}: {statistics.mean(ms)*1e3:8.1f} us  -> {statistics.mean(ms)*1e3/24:6.2f} us per batch-iteration, frac {3*167.1168e6*G/(statistics.mean(ms)*1e-3)/1e9/6545.3:.4f}", flush=True)
