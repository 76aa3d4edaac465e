"""Small pyramid levels on the resident variant of the iteration kernel (uic_iter_kernel<.., RES>): the live frame of a
pair is copied to shared memory and the footprint is looked up there.  Only the way the taps are fetched changes, so
the masks must equal those of the global-memory lookups (options.small_levels = 1) bit for bit, the sums to fp32
summation order, and everything stays within the stated tolerances of the oracle."""
import pytest
import torch

from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import _twist_to_pose, levels_to, make_frame_pairs
from helpers import TOL_POSE, TOL_SYS, frob_rel
from oracle import ic_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
PLAIN = dict(small_levels=1)


def same(a, b):
    """Sums of the two variants: the arithmetic per pixel is identical, the tile heights (hence the order of the
    fp32 partial sums) may differ."""
    return frob_rel(a, b) < 2e-6


def solve_both(levels, pose, **kw):
    a = A.uic_solve(levels, pose, **kw)
    b = A.uic_solve(levels, pose, tuning=PLAIN, **kw)
    torch.cuda.synchronize()
    return a, b


@pytest.mark.parametrize("B,C,H,W,tru", [(5, 8, 30, 40, True), (5, 8, 15, 20, True), (3, 8, 30, 40, False),
                                         (4, 3, 15, 21, True),      # planes that are no multiple of 16 bytes: 4-byte copies
                                         (2, 4, 37, 23, True), (7, 2, 9, 70, False), (3, 16, 24, 32, True)])
def test_resident_equals_global_lookups_and_the_oracle(B, C, H, W, tru):
    data = make_frame_pairs(B, C, H, W, seed=100 + H + W, n_levels=1)
    lv = data["levels"][0]
    assert lv["s0"].shape[1] == C       # the uncertainty repeated to C channels, as the modules receive it
    if (H + W) % 2 == 1:                # half of the cases with the ONE map the encoder emits
        lv = dict(lv, s0=lv["s0"][:, :1].contiguous(), s1=lv["s1"][:, :1].contiguous())
    pose = (data["R0"], data["t0"])
    a, b = solve_both(levels_to([lv], DEV), (pose[0].to(DEV), pose[1].to(DEV)), iters=3, remove_tru_sigma=tru, want_occ=True)
    assert torch.equal(a.occ[0], b.occ[0])
    assert same(a.sys_hist, b.sys_hist) and (a.pose_hist - b.pose_hist).abs().max() < 1e-6
    trace = []
    s0, s1 = lv["s0"].expand(-1, C, -1, -1), lv["s1"].expand(-1, C, -1, -1)
    (R, t), _ = O.uic_level(pose, lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"], s0, s1, iters=3,
                            remove_tru_sigma=tru, trace=trace)
    Ac, bc = A.unpack_system(a.sys_hist[0].cpu())
    assert frob_rel(Ac, trace[0]["A"]) < TOL_SYS and frob_rel(bc, trace[0]["b"]) < 5 * TOL_SYS
    assert torch.equal(a.occ[0][0].cpu(), trace[0]["occ"][:, 0].to(torch.uint8))
    assert (a.pose[0].cpu() - R).abs().max() < TOL_POSE and (a.pose[1].cpu() - t).abs().max() < TOL_POSE


def test_large_motion_and_depth_holes():
    """Lookups anywhere in the frame (70 degrees in-plane rotation, a third of the depth missing): the resident copy
    holds the whole frame, so nothing depends on where the footprints fall."""
    data = make_frame_pairs(3, 8, 30, 40, seed=18, n_levels=1)
    lv = data["levels"][0]
    hole = torch.rand(lv["invD0"].shape, generator=torch.Generator().manual_seed(2)) < 0.3
    lv["invD0"] = torch.where(hole, torch.zeros_like(lv["invD0"]), lv["invD0"]).contiguous()
    pose = _twist_to_pose(torch.tensor([[0.0, 0.35, 0.0, 0.4, 0.1, 0.0], [0.2, 0.0, 0.3, -0.3, 0.2, 0.1],
                                        [0.0, 0.0, 1.2, 0.0, 0.0, 0.0]]))
    a, b = solve_both(levels_to([lv], DEV), (pose[0].to(DEV), pose[1].to(DEV)), iters=2, remove_tru_sigma=True, want_occ=True)
    assert torch.equal(a.occ[0], b.occ[0]) and same(a.sys_hist, b.sys_hist)
    trace = []
    O.uic_level(pose, lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"], lv["s0"].expand(-1, 8, -1, -1),
                lv["s1"].expand(-1, 8, -1, -1), iters=2, remove_tru_sigma=True, trace=trace)
    assert torch.equal(a.occ[0][0].cpu(), trace[0]["occ"][:, 0].to(torch.uint8))
    assert trace[0]["occ"].float().mean() > 0.3


def test_object_masks_single_sigma_map_and_shared_keyframe():
    B, C, H, W = 6, 8, 30, 40
    data = make_frame_pairs(B, C, H, W, seed=7, n_levels=1)
    lv = levels_to(data["levels"], DEV)[0]
    pose = (data["R0"].to(DEV), data["t0"].to(DEV))
    gen = torch.Generator().manual_seed(3)
    m0 = (torch.rand((B, 1, H, W), generator=gen) > 0.2).to(DEV)
    m1 = (torch.rand((B, 1, H, W), generator=gen) > 0.2).to(DEV)
    a, b = solve_both([lv], pose, iters=3, remove_tru_sigma=True, want_occ=True, obj_mask0=[m0], obj_mask1=[m1])
    assert torch.equal(a.occ[0], b.occ[0]) and same(a.sys_hist, b.sys_hist)
    # one uncertainty map per frame (what the reference's encoder emits) against the repeated tensor
    one = dict(lv, s0=lv["s0"][:, :1].contiguous(), s1=lv["s1"][:, :1].contiguous())
    rep = dict(lv, s0=one["s0"].repeat(1, C, 1, 1), s1=one["s1"].repeat(1, C, 1, 1))
    c, d = solve_both([one], pose, iters=3, remove_tru_sigma=True)
    e = A.uic_solve([rep], pose, iters=3, remove_tru_sigma=True)
    assert same(c.sys_hist, d.sys_hist)
    assert (c.pose_hist - e.pose_hist).abs().max() < 1e-6
    # one keyframe for the whole batch (kf_vo tracking)
    kf = dict(lv, x0=lv["x0"][:1].contiguous(), s0=lv["s0"][:1].contiguous(), invD0=lv["invD0"][:1].contiguous())
    f, g = solve_both([kf], pose, iters=3, remove_tru_sigma=True, shared_keyframe=True)
    assert same(f.sys_hist, g.sys_hist)


def test_bench_pyramid_in_groups():
    """The coarse levels of the bench configuration (15x20, 30x40 resident; 60x80, 120x160 staged), three batches in one
    call with their own sigma extremes: equal to the global-lookup run, and to separate calls."""
    B, G = 8, 3
    parts = [make_frame_pairs(B, 8, 120, 160, seed=50 + g, n_levels=4) for g in range(G)]
    levels = [{k: torch.cat([p["levels"][l][k] for p in parts]).to(DEV) for k in parts[0]["levels"][l]} for l in range(4)]
    pose = (torch.cat([p["R0"] for p in parts]).to(DEV), torch.cat([p["t0"] for p in parts]).to(DEV))
    a, b = solve_both(levels, pose, iters=3, remove_tru_sigma=True, group=B)
    assert same(a.sys_hist[:6], b.sys_hist[:6])               # the two resident levels
    assert (a.pose_hist - b.pose_hist).abs().max() < 1e-6
    for g in range(G):
        one = A.uic_solve(levels_to(parts[g]["levels"], DEV), (parts[g]["R0"].to(DEV), parts[g]["t0"].to(DEV)), iters=3,
                          remove_tru_sigma=True)
        assert (a.pose[1][g * B:(g + 1) * B] - one.pose[1]).abs().max() < 1e-6
