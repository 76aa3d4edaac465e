"""Full (B,C,H,W) uncertainty tensors whose channels are copies of channel 0 -- what the reference's encoder hands to
the tracker (algorithms.py:1425-1427) -- are found by a device-side check and served by the one-map tile routines
(dpft_uic_options.sigma_detect).  The results must be those of the C-map routines; tensors that are NOT replicated
(even in one element) must take the C-map routines; and because the decision is made on the device, a CUDA graph of
the call stays right when its buffers are refilled with the other kind of data."""
import pytest
import torch

from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.batched import BatchedSolver
from deep_prob_feature_track_b200.synthetic import levels_to, make_frame_pairs
from helpers import TOL_POSE, frob_rel
from oracle import ic_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
OFF = dict(sigma_detect=1)


def stacked(B, G, seed, sigma_channels=1, H=120, W=160, n_levels=4):
    parts = [make_frame_pairs(B, 8, H, W, seed=seed + g, n_levels=n_levels, sigma_channels=sigma_channels) for g in range(G)]
    levels = [{k: torch.cat([p["levels"][l][k] for p in parts]).to(DEV) for k in parts[0]["levels"][l]} for l in range(n_levels)]
    for lv in levels:
        lv["s0"] = lv["s0"].expand(-1, 8, -1, -1).contiguous()
        lv["s1"] = lv["s1"].expand(-1, 8, -1, -1).contiguous()
    pose = (torch.cat([p["R0"] for p in parts]).to(DEV), torch.cat([p["t0"] for p in parts]).to(DEV))
    return parts, levels, pose


def check_sigma0_extremes(res, levels, B, iters=3):
    """aux_hist[k, g, 2:4] = min / max of sigma0 of the level of iteration k over the pairs of group g (alg:1976-1977).  With
    the check on, the replication kernel collects them on its way; they must be torch's, bit for bit."""
    for l, lv in enumerate(levels):
        s0 = lv["s0"]
        G = s0.shape[0] // B
        lo = s0.reshape(G, -1).amin(dim=1)
        hi = s0.reshape(G, -1).amax(dim=1)
        for it in range(iters):
            a = res.aux_hist[l * iters + it]
            assert torch.equal(a[:, 2], lo) and torch.equal(a[:, 3], hi), (l, it)


@pytest.mark.parametrize("queue", [False, True])
def test_replicated_tensors_equal_the_c_map_routines_and_the_one_map_call(queue):
    B, G = 8, 3
    parts, levels, pose = stacked(B, G, 300)
    kw = dict(iters=3, remove_tru_sigma=True, group=B, queue=queue)
    a = A.uic_solve(levels, pose, **kw)                                   # detection on (default)
    b = A.uic_solve(levels, pose, tuning=OFF, **kw)                       # C-map routines
    one = [dict(lv, s0=lv["s0"][:, :1].contiguous(), s1=lv["s1"][:, :1].contiguous()) for lv in levels]
    c = A.uic_solve(one, pose, **kw)                                      # the caller passes the one map itself
    torch.cuda.synchronize()
    a.raise_if_bad()
    assert (a.pose_hist - b.pose_hist).abs().max() < 1e-6
    assert frob_rel(a.sys_hist, b.sys_hist) < 2e-6
    assert (a.pose_hist - c.pose_hist).abs().max() < 1e-6 and frob_rel(a.sys_hist, c.sys_hist) < 2e-6
    check_sigma0_extremes(a, levels, B)
    check_sigma0_extremes(b, levels, B)
    # ... and the oracle, first batch
    trace = []
    with torch.no_grad():
        (R, t), _ = O.track_pyramid(parts[0]["levels"], (parts[0]["R0"], parts[0]["t0"]), iters=3, remove_tru_sigma=True,
                                    trace=trace)
    Rc, tc = a.pose
    assert (Rc[:B].cpu() - R).abs().max() < TOL_POSE and (tc[:B].cpu() - t).abs().max() < TOL_POSE


@pytest.mark.parametrize("queue", [False, True])
def test_tensors_that_are_not_replicated_take_the_c_map_routines(queue):
    B, G = 8, 2
    _, levels, pose = stacked(B, G, 320, sigma_channels=8)
    assert not torch.equal(levels[-1]["s1"][:, 0], levels[-1]["s1"][:, 7])
    kw = dict(iters=3, remove_tru_sigma=True, group=B, queue=queue)
    a = A.uic_solve(levels, pose, **kw)
    b = A.uic_solve(levels, pose, tuning=OFF, **kw)
    torch.cuda.synchronize()
    assert torch.equal(a.pose_hist, b.pose_hist) and torch.equal(a.sys_hist, b.sys_hist)
    check_sigma0_extremes(a, levels, B)     # extremes over ALL channels, collected by the check although it found differences


@pytest.mark.parametrize("where", ["s1_finest_last", "s0_coarsest_first", "s1_level1_middle"])
def test_one_differing_element_is_enough(where):
    B, G = 8, 2
    _, levels, pose = stacked(B, G, 340)
    lv = {"s1_finest_last": levels[3], "s0_coarsest_first": levels[0], "s1_level1_middle": levels[2]}[where]
    if where == "s1_finest_last":
        lv["s1"][-1, 7, -1, -1] *= 1.0 + 2 ** -20
    elif where == "s0_coarsest_first":
        lv["s0"][0, 1, 0, 0] *= 1.0 + 2 ** -20
    else:
        lv["s1"][B, 3, 17, 41] += 0.25
    kw = dict(iters=3, remove_tru_sigma=True, group=B, queue=True)
    a = A.uic_solve(levels, pose, **kw)
    b = A.uic_solve(levels, pose, tuning=OFF, **kw)
    torch.cuda.synchronize()
    assert torch.equal(a.pose_hist, b.pose_hist) and torch.equal(a.sys_hist, b.sys_hist)


def test_shared_keyframe_and_small_single_level():
    data = make_frame_pairs(6, 8, 30, 40, seed=9, n_levels=1)
    lv = levels_to(data["levels"], DEV)[0]
    lv = dict(lv, s0=lv["s0"].expand(-1, 8, -1, -1).contiguous(), s1=lv["s1"].expand(-1, 8, -1, -1).contiguous())
    pose = (data["R0"].to(DEV), data["t0"].to(DEV))
    a = A.uic_solve([lv], pose, iters=3, remove_tru_sigma=True)
    b = A.uic_solve([lv], pose, iters=3, remove_tru_sigma=True, tuning=OFF)
    assert (a.pose_hist - b.pose_hist).abs().max() < 1e-6 and frob_rel(a.sys_hist, b.sys_hist) < 2e-6
    kf = dict(lv, x0=lv["x0"][:1].contiguous(), s0=lv["s0"][:1].contiguous(), invD0=lv["invD0"][:1].contiguous())
    c = A.uic_solve([kf], pose, iters=3, remove_tru_sigma=True, shared_keyframe=True)
    d = A.uic_solve([kf], pose, iters=3, remove_tru_sigma=True, shared_keyframe=True, tuning=OFF)
    torch.cuda.synchronize()
    assert (c.pose_hist - d.pose_hist).abs().max() < 1e-6 and frob_rel(c.sys_hist, d.sys_hist) < 2e-6


def test_a_captured_graph_follows_the_data():
    """Capture with replicated uncertainty, refill the same buffers with independent channels (and back): every replay
    gives what a plain call on that data gives."""
    B, G = 8, 2
    _, rep, pose = stacked(B, G, 360)
    _, ind, _ = stacked(B, G, 380, sigma_channels=8)
    want_rep = A.uic_solve(rep, pose, iters=3, remove_tru_sigma=True, group=B, queue=True, tuning=OFF)
    want_ind = A.uic_solve(ind, pose, iters=3, remove_tru_sigma=True, group=B, queue=True, tuning=OFF)
    solver = BatchedSolver(B, iters=3, remove_tru_sigma=True, streams=1, device=torch.device(DEV), graphs=True, queue=True)
    buf = [{k: v.clone() for k, v in lv.items()} for lv in rep]

    def refill(src):
        for lv, s in zip(buf, src):
            for k in lv:
                lv[k].copy_(s[k])

    r = solver.submit(buf, pose)
    solver.synchronize()
    assert (r.pose_hist - want_rep.pose_hist).abs().max() < 1e-6
    refill(ind)
    r = solver.submit(buf, pose)
    solver.synchronize()
    assert solver.replays >= 2
    assert torch.equal(r.pose_hist, want_ind.pose_hist) and torch.equal(r.sys_hist, want_ind.sys_hist)
    refill(rep)
    r = solver.submit(buf, pose)
    solver.synchronize()
    r.raise_if_bad()
    assert (r.pose_hist - want_rep.pose_hist).abs().max() < 1e-6
