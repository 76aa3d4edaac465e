"""Keyframe-mode tracking (one keyframe, many live frames in one call) against B = 1 oracle runs per frame."""
import pytest
import torch

from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import levels_to, make_frame_pairs
from helpers import TOL_POSE, TOL_SYS, frob_rel
from oracle import ic_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_batched_keyframe_tracking_equals_per_frame_reference_calls():
    n_live, C, H, W = 6, 8, 48, 64
    data = make_frame_pairs(n_live, C, H, W, seed=77, n_levels=3)
    # one keyframe (pair 0's) for every live frame
    key = [{k: lv[k][:1].contiguous() for k in ("x0", "s0", "invD0")} for lv in data["levels"]]
    live = [{k: lv[k] for k in ("x1", "s1", "invD1", "K")} for lv in data["levels"]]
    tracker = A.KeyframeTracker(levels_to(key, DEV), iters=3, remove_tru_sigma=True)
    res = tracker.track(levels_to(live, DEV), (data["R0"].to(DEV), data["t0"].to(DEV)))
    torch.cuda.synchronize()
    assert int(res.status.item()) == 0
    R, t = (x.cpu() for x in res.pose)
    for i in range(n_live):
        lv_i = [dict(x0=k["x0"], s0=k["s0"], invD0=k["invD0"], x1=l["x1"][i:i + 1], s1=l["s1"][i:i + 1],
                     invD1=l["invD1"][i:i + 1], K=l["K"][i:i + 1]) for k, l in zip(key, live)]
        trace = []
        with torch.no_grad():
            (Ro, to), _ = O.track_pyramid(lv_i, (data["R0"][i:i + 1], data["t0"][i:i + 1]), iters=3,
                                          remove_tru_sigma=True, trace=trace)
        assert (R[i] - Ro[0]).abs().max() < TOL_POSE and (t[i] - to[0]).abs().max() < TOL_POSE
        A0, b0 = A.unpack_system(res.sys_hist[0, i].cpu())
        assert frob_rel(A0, trace[0][0]["A"][0]) < TOL_SYS and frob_rel(b0, trace[0][0]["b"][0]) < TOL_SYS


def test_pairwise_extremes_without_shared_keyframe():
    B, C, H, W = 4, 4, 30, 40
    data = make_frame_pairs(B, C, H, W, seed=5, n_levels=1)
    lv = data["levels"][0]
    res = A.uic_solve(levels_to([lv], DEV), (data["R0"].to(DEV), data["t0"].to(DEV)), iters=2, remove_tru_sigma=True,
                      pairwise_extremes=True)
    R, t = (x.cpu() for x in res.pose)
    for i in range(B):
        one = {k: v[i:i + 1] for k, v in lv.items()}
        (Ro, to), _ = O.uic_level((data["R0"][i:i + 1], data["t0"][i:i + 1]), one["x0"], one["x1"], one["invD0"],
                                  one["invD1"], one["K"], one["s0"], one["s1"], iters=2, remove_tru_sigma=True)
        assert (R[i] - Ro[0]).abs().max() < TOL_POSE and (t[i] - to[0]).abs().max() < TOL_POSE
