"""Diagnostic for narrow_probe's one outlier (G=12, queue_levels=4 narrow: |pose| 2.5e-4 against the default policy):
which pair / iteration deviates first, and how both solves compare with the oracle there."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import make_frame_pairs
from oracle import ic_oracle as O

dev = torch.device("cuda:0")
B, C, H, W = 64, 8, 120, 160
G = int(sys.argv[1]) if len(sys.argv) > 1 else 12
parts = [make_frame_pairs(B, C, H, W, seed=1234 + g, n_levels=4) for g in range(G)]
levels_cpu = [{k: torch.cat([p["levels"][l][k] for p in parts]) for k in parts[0]["levels"][l]} for l in range(4)]
levels = [{k: v.to(dev) for k, v in lv.items()} for lv in levels_cpu]
for lv in levels:
    lv["s0"] = lv["s0"].expand(-1, C, -1, -1).contiguous(); lv["s1"] = lv["s1"].expand(-1, C, -1, -1).contiguous()
pose_cpu = (torch.cat([p["R0"] for p in parts]), torch.cat([p["t0"] for p in parts]))
pose = (pose_cpu[0].to(dev), pose_cpu[1].to(dev))
kw = dict(iters=3, remove_tru_sigma=True, group=B)
runs = {"default": A.uic_solve(levels, pose, **kw),
        "ql4 narrow": A.uic_solve(levels, pose, queue_levels=4, **kw),
        "ql4 narrow again": A.uic_solve(levels, pose, queue_levels=4, **kw),
        "ql4 plain": A.uic_solve(levels, pose, queue_levels=4, tuning=dict(small_levels=2), **kw)}
torch.cuda.synchronize()
ref = runs["default"]
print("bitwise repeatable:", torch.equal(runs["ql4 narrow"].pose_hist, runs["ql4 narrow again"].pose_hist))
for name, r in runs.items():
    d = (r.pose_hist - ref.pose_hist).abs().amax(dim=2)       # (n_it + 1, B)
    bad = (d > 1e-6).nonzero()
    print(f"{name}: max |pose diff| {d.max().item():.2e}; entries > 1e-6: {bad.shape[0]}")
    if bad.shape[0]:
        k, b = int(bad[0, 0]), int(bad[0, 1])
        print(f"   first: pose row {k} (written by iteration {k - 1}), pair {b} (batch {b // B}, index {b % B})")
        kk = k - 1
        sa, sb = r.sys_hist[kk, b].cpu(), ref.sys_hist[kk, b].cpu()
        print(f"   sys of iteration {kk}: rel diff {((sa - sb).abs().max() / sb.abs().max()).item():.2e}; pose entering it differs by {d[kk, b].item():.2e}")
        g0 = (b // B) * B
        lv_b = [{k2: v[g0:g0 + B] for k2, v in lv.items()} for lv in levels_cpu]
        trace = []
        with torch.no_grad():
            O.track_pyramid(lv_b, (pose_cpu[0][g0:g0 + B], pose_cpu[1][g0:g0 + B]), iters=3, remove_tru_sigma=True, trace=trace, reduction="einsum")
        # the oracle started from THIS run's own poses entering iteration kk (a mask flip of a threshold-adjacent pixel
        # caused by last-bit differences in the pose would reproduce here; a wrong lookup would not)
        lvl = kk // 3
        Rn, tn = A.unpack_pose(r.pose_hist[kk, g0:g0 + B].cpu())
        tr2 = []
        with torch.no_grad():
            q = lv_b[lvl]
            O.uic_level((Rn, tn), q["x0"], q["x1"], q["invD0"], q["invD1"], q["K"], q["s0"], q["s1"], iters=1,
                        remove_tru_sigma=True, reduction="einsum", trace=tr2)
        Ac, bc = A.unpack_system(r.sys_hist[kk, g0:g0 + B].cpu())
        i = b % B
        print(f"   {name}: against the oracle restarted from its own pose: A {((Ac[i] - tr2[0]['A'][i]).norm() / tr2[0]['A'][i].norm()).item():.2e}, "
              f"b {((bc[i] - tr2[0]['b'][i]).norm() / tr2[0]['b'][i].norm()).item():.2e}; occ pixels of the pair: oracle restarted "
              f"{int(tr2[0]['occ'][i].sum())}, oracle chain {int(trace[lvl][kk % 3]['occ'][i].sum())}")
        for nm, rr in (("default", ref), (name, r)):
            for it in range(kk, min(kk + 3, 12)):
                rec = trace[it // 3][it % 3]
                Ac, bc = A.unpack_system(rr.sys_hist[it, g0:g0 + B].cpu())
                i = b % B
                ea = ((Ac[i] - rec["A"][i]).norm() / rec["A"][i].norm()).item()
                eb = ((bc[i] - rec["b"][i]).norm() / rec["b"][i].norm()).item()
                Rc, tc = A.unpack_pose(rr.pose_hist[it + 1, g0:g0 + B].cpu())
                Ro, to = trace[it // 3][it % 3]["R"], trace[it // 3][it % 3]["t"]
                print(f"   {nm:18s} it {it}: A rel err vs oracle {ea:.2e}, b {eb:.2e}")
