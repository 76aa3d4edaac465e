import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import torch
from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import levels_to, make_frame_pairs
from helpers import frob_rel
from oracle import ic_oracle as O
DEV = "cuda:0"
B, C, H, W = 6, 8, 40, 64
data = make_frame_pairs(B, C, H, W, seed=77, n_levels=1)
lv = data["levels"][0]
for k in ("s0", "s1"):
    lv[k] = lv[k].clamp(0.8, 1.25).contiguous()
pose0 = (data["R0"], data["t0"])
trace = []
O.uic_level(pose0, lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"], lv["s0"], lv["s1"], iters=1, remove_tru_sigma=True, trace=trace)
def run(**kw):
    r = A.uic_solve(levels_to([lv], DEV), (pose0[0].to(DEV), pose0[1].to(DEV)), iters=1, remove_tru_sigma=True, **kw)
    torch.cuda.synchronize()
    Ac, bc = A.unpack_system(r.sys_hist[0].cpu())
    return frob_rel(Ac, trace[0]["A"]), frob_rel(bc, trace[0]["b"]), r
for name, kw in (("lpi staged", dict(queue=False)), ("lpi plain", dict(queue=False, staged_footprint=False)), ("lpi occ", dict(queue=False, want_occ=True)),
                 ("queue tr7", dict(queue=True, tile_rows=[7])), ("queue tr40", dict(queue=True, tile_rows=[40])), ("queue plain", dict(queue=True, staged_footprint=False)),
                 ("queue pairwise", dict(queue=True, group=1)), ("lpi pairwise", dict(queue=False, group=1)), ("queue 1cta", dict(queue=True, queue_ctas=1))):
    a, b, r = run(**kw)
    print(f"{name:16s} relA {a:.2e} relb {b:.2e} aux {r.aux_hist.reshape(-1, 4)[0].tolist()}")
sr = None
print("oracle occ frac", trace[0]["occ"].float().mean().item())
