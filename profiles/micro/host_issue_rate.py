import sys, time, torch
sys.path.insert(0, '/root/repo')
from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import levels_to, make_frame_pairs
data = make_frame_pairs(2, 8, 16, 24, seed=1, n_levels=2)
lv = levels_to(data["levels"], "cuda:0"); pose = (data["R0"].cuda(), data["t0"].cuda())
for _ in range(20): A.uic_solve(lv, pose, iters=3, remove_tru_sigma=True)
torch.cuda.synchronize(); t=time.perf_counter()
n=500
for _ in range(n): A.uic_solve(lv, pose, iters=3, remove_tru_sigma=True)
t1=time.perf_counter()-t; torch.cuda.synchronize(); t2=time.perf_counter()-t
print("host issue time per solve (2 levels): %.1f us, incl. drain %.1f us" % (t1/n*1e6, t2/n*1e6))
data = make_frame_pairs(2, 8, 32, 48, seed=1, n_levels=4)
lv = levels_to(data["levels"], "cuda:0")
for _ in range(20): A.uic_solve(lv, pose, iters=3, remove_tru_sigma=True)
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(n): A.uic_solve(lv, pose, iters=3, remove_tru_sigma=True)
t1=time.perf_counter()-t; torch.cuda.synchronize(); t2=time.perf_counter()-t
print("host issue time per solve (4 levels): %.1f us, incl. drain %.1f us" % (t1/n*1e6, t2/n*1e6))
