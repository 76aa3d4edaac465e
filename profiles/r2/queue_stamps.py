"""Phase stamps of pair 0 inside the work-queue launch (library built with DPFT_NVCC_EXTRA=-DDPFT_QUEUE_STAMPS)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from deep_prob_feature_track_b200 import _lib, algorithms as A
from deep_prob_feature_track_b200.synthetic import make_frame_pairs

dev = torch.device("cuda:0")
B, C, H, W = 64, 8, 120, 160
d = make_frame_pairs(B, C, H, W, seed=1234, n_levels=4)
levels = [{k: v.to(dev) for k, v in lv.items()} for lv in d["levels"]]
pose = (d["R0"].to(dev), d["t0"].to(dev))
L = _lib.lib()
L.dpft_debug_queue_stamps.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
for i in range(3):
    A.uic_solve(levels, pose, iters=3, remove_tru_sigma=True)
torch.cuda.synchronize()
L.dpft_debug_queue_stamps(None, 0, 1)
A.uic_solve(levels, pose, iters=3, remove_tru_sigma=True)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (16 * 12))()
L.dpft_debug_queue_stamps(buf, 12, 0)
t0 = buf[0]
names = ["first tile dequeued", "last tile: walk start", "walk end", "flushed", "counted", "folded", "solved/parked", "pushed", "group resolved"]
for k in range(12):
    row = [buf[k * 16 + i] for i in range(9)]
    print(f"k={k:2d} " + "  ".join(f"{n.split()[0]}={(v - t0) / 1e3:8.1f}" if v != 2**64 - 1 else f"{n.split()[0]}=   -    " for n, v in zip(names, row)))
