"""One small call of every entry point, as a target for ncu / memory checkers (staged / plain / single-sigma forward, backward,
ICP, IC tracker with its backward, residual loss, depth stage, pose loss)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deep_prob_feature_track_b200 import algorithms as A, criterions as C
from deep_prob_feature_track_b200.synthetic import levels_to, make_frame_pairs

dev = "cuda:0"
data = make_frame_pairs(3, 8, 60, 80, seed=5, n_levels=3, with_depth=True)
lv = levels_to(data["levels"], dev)
pose = (data["R0"].to(dev), data["t0"].to(dev))
for kw in (dict(), dict(staged_footprint=False), dict(single_launch=True), dict(fused_sobel=False), dict(combine_icp=True)):
    A.uic_solve(lv, pose, iters=2, remove_tru_sigma=True, want_occ=not kw.get("single_launch", False), **kw).raise_if_bad()
one = [dict(l, s0=l["s0"][:, :1].contiguous(), s1=l["s1"][:, :1].contiguous()) for l in lv]
A.uic_solve(one, pose, iters=2, remove_tru_sigma=True).raise_if_bad()
A.uic_residual_loss(lv[-1], pose, remove_tru_sigma=True)
key = [{k: l[k][:1].contiguous() for k in ("x0", "s0", "invD0")} for l in lv]
A.KeyframeTracker(key, iters=2).track([{k: l[k] for k in ("x1", "s1", "invD1", "K")} for l in lv], pose)
g = [dict(l) for l in lv]
for l in g:
    for k in ("x0", "x1", "s0", "s1"):
        l[k] = l[k].clone().requires_grad_(True)
outs = A.uic_track(g, pose, iters=2, remove_tru_sigma=True, combine_icp=True)
sum((R.sum() + t.sum()) for R, t, _ in outs).backward()
d1 = make_frame_pairs(2, 1, 24, 32, seed=6, n_levels=1)["levels"][0]
t = {k: v.to(dev) for k, v in d1.items()}
x0, x1 = t["x0"].clone().requires_grad_(True), t["x1"].clone().requires_grad_(True)
mod = A.TrustRegionBase(max_iter=2, mEst_func=None, solver_func=A.DirectSolverNet("Direct-ResVol")).to(dev)
(R, tt), _ = mod([torch.eye(3, device=dev).repeat(2, 1, 1), torch.zeros(2, 3, device=dev)], x0, x1, t["invD0"], t["invD1"], t["K"])
(R.sum() + tt.sum()).backward()
depth = torch.rand(2, 1, 120, 160, device=dev) + 0.5
A.depth_pyramids(depth, 4, with_depth=True)
Re = torch.eye(3, device=dev).repeat(2, 4, 1, 1).requires_grad_(True)
te = (torch.rand(2, 4, 3, device=dev) * 0.1).requires_grad_(True)
loss = C.compute_RT_EPE_loss(Re, te, torch.eye(3, device=dev).repeat(2, 1, 1), torch.zeros(2, 3, device=dev), depth,
                             torch.tensor([[131.0, 131.0, 80.0, 60.0]], device=dev).repeat(2, 1),
                             invalid=torch.zeros(2, 1, 120, 160, device=dev, dtype=torch.bool))
loss.sum().backward()
torch.cuda.synchronize()
print("ok")
