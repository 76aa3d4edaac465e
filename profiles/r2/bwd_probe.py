"""Round 2 probe: solver forward + backward (autograd through 4 levels x 3 iterations, B = 64, 120x160, C = 8), used to
compare builds of the backward (DPFT_LIB_PATH=profiles/r2/variants/NAME.so).

Recorded experiment (r2b_bwd_probe.txt): a tile-structured backward (warp tiles as in the forward, unit Sobel recomputed,
its adjoint gathered in registers, no materialised gradient maps) moved 2.2x fewer bytes but ran 2.8 ms against 1.9 ms
for the one-thread-per-pixel kernel: a serial row loop per warp at 12 warps per SM exposes every load and reduction
latency that the pixel kernel hides behind 16 warps of short independent threads.  Not kept."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import make_frame_pairs

dev = torch.device("cuda:0")
B, C, H, W = 64, 8, 120, 160
data = make_frame_pairs(B, C, H, W, seed=1234, n_levels=4)
levels = [{k: v.to(dev) for k, v in lv.items()} for lv in data["levels"]]
leaves = [{k: (v.clone().requires_grad_(True) if k in ("x0", "x1", "s0", "s1") else v) for k, v in lv.items()} for lv in levels]
R = data["R0"].to(dev).clone().requires_grad_(True)
t = data["t0"].to(dev).clone().requires_grad_(True)
gen = torch.Generator().manual_seed(1)
cs = [(torch.randn((B, 3, 3), generator=gen).to(dev), torch.randn((B, 3), generator=gen).to(dev)) for _ in range(4)]


def step(tiled, fwd_only=False):
    for lv in leaves:
        for k in ("x0", "x1", "s0", "s1"):
            lv[k].grad = None
    R.grad = t.grad = None
    outs = A.uic_track(leaves, (R, t), iters=3, remove_tru_sigma=True, check=False)
    if fwd_only:
        return
    sum((Rl * c[0]).sum() + (tl * c[1]).sum() for (Rl, tl, _), c in zip(outs, cs)).backward()


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


fwd = timeit(lambda: step(True, fwd_only=True))
ms = timeit(lambda: step(True))
print(f"{os.environ.get('DPFT_LIB_PATH', 'default build')}: forward+backward {ms:.3f} ms per batch of {B} (forward alone {fwd:.3f} ms)", flush=True)
