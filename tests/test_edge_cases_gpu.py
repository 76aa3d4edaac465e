"""Edge cases of the solver path against the oracle: degenerate sizes, fully masked frames, holes in the
depth, out-of-view motion, non-contiguous inputs, larger batches, and the error behaviour."""
import pytest
import torch

from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import _twist_to_pose, levels_to, make_frame_pairs
from helpers import TOL_POSE, TOL_SYS, frob_rel
from oracle import ic_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def both(lv, pose, iters=2, tru=True, **kw):
    res = A.uic_solve(levels_to([lv], DEV), (pose[0].to(DEV), pose[1].to(DEV)), iters=iters, remove_tru_sigma=tru,
                      want_occ=True, **kw)
    trace = []
    O.uic_level(pose, lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"], lv["s0"], lv["s1"], iters=iters,
                remove_tru_sigma=tru, trace=trace)
    torch.cuda.synchronize()
    return res, trace


def check(res, trace, mask_slack=0):
    for it, rec in enumerate(trace):
        Ac, bc = A.unpack_system(res.sys_hist[it].cpu())
        assert frob_rel(Ac, rec["A"]) < TOL_SYS
        assert frob_rel(bc, rec["b"]) < 5 * TOL_SYS
        flips = int((res.occ[0][it].cpu() != rec["occ"][:, 0].to(torch.uint8)).sum())
        assert flips <= (0 if it == 0 else mask_slack), (it, flips)


@pytest.mark.parametrize("H,W", [(2, 2), (3, 5), (2, 64), (33, 2), (5, 95)])
def test_tiny_and_thin_images(H, W):
    data = make_frame_pairs(2, 4, H, W, seed=H * 100 + W, n_levels=1)
    pose = _twist_to_pose(torch.tensor([[0.01, -0.02, 0.005, 0.01, 0.0, -0.01]] * 2))
    res, trace = both(data["levels"][0], pose, iters=1)
    check(res, trace)


def test_everything_masked():
    """Inverse depths that never agree: every pixel is excluded, J^T r is the sum of J * 1e-6 (alg:1982-1983)."""
    data = make_frame_pairs(2, 4, 20, 28, seed=3, n_levels=1)
    lv = data["levels"][0]
    lv["invD1"] = (lv["invD1"] + 5.0).contiguous()
    res, trace = both(lv, (data["R0"], data["t0"]), iters=1, tru=False)
    assert bool(trace[0]["occ"].all())
    check(res, trace)


def test_depth_holes_and_out_of_view_motion():
    data = make_frame_pairs(3, 8, 30, 40, seed=8, n_levels=1)
    lv = data["levels"][0]
    hole = torch.rand(lv["invD0"].shape, generator=torch.Generator().manual_seed(1)) < 0.3
    lv["invD0"] = torch.where(hole, torch.zeros_like(lv["invD0"]), lv["invD0"]).contiguous()   # invalid depth = 0
    pose = _twist_to_pose(torch.tensor([[0.0, 0.35, 0.0, 0.4, 0.1, 0.0]] * 3))                 # most pixels leave the view
    res, trace = both(lv, pose, iters=2)
    assert trace[0]["occ"].float().mean() > 0.4
    check(res, trace, mask_slack=3)


def test_large_batch_and_many_channels():
    data = make_frame_pairs(150, 16, 15, 20, seed=21, n_levels=1)
    gen = torch.Generator().manual_seed(4)
    pose = _twist_to_pose((torch.rand((150, 6), generator=gen) * 2 - 1) * 0.02)
    res, trace = both(data["levels"][0], pose, iters=2)
    check(res, trace, mask_slack=2)
    R, t = res.pose
    (Ro, to) = O.uic_level(pose, *(data["levels"][0][k] for k in ("x0", "x1", "invD0", "invD1", "K", "s0", "s1")),
                           iters=2, remove_tru_sigma=True)[0]
    assert (R.cpu() - Ro).abs().max() < TOL_POSE and (t.cpu() - to).abs().max() < TOL_POSE


def test_non_contiguous_and_double_inputs_are_accepted():
    data = make_frame_pairs(2, 4, 24, 32, seed=5, n_levels=1)
    lv = {k: v.to(DEV) for k, v in data["levels"][0].items()}
    ref = A.uic_solve([lv], (data["R0"].to(DEV), data["t0"].to(DEV)), iters=2, remove_tru_sigma=True)
    odd = dict(lv)
    odd["x0"] = lv["x0"].permute(0, 1, 3, 2).contiguous().permute(0, 1, 3, 2)   # same values, transposed strides
    odd["s1"] = lv["s1"].double()
    out = A.uic_solve([odd], (data["R0"].to(DEV).double(), data["t0"].to(DEV).view(2, 3, 1)), iters=2,
                      remove_tru_sigma=True)
    assert torch.equal(out.pose_hist, ref.pose_hist)


def test_non_finite_input_raises_like_the_reference():
    data = make_frame_pairs(2, 4, 16, 20, seed=6, n_levels=1)
    lv = {k: v.to(DEV) for k, v in data["levels"][0].items()}
    lv["x1"] = lv["x1"].clone()
    lv["x1"][0, 1, 5, 7] = float("nan")
    mod = A.TrustRegionInverseWUncertainty(max_iter=1)
    with torch.no_grad(), pytest.raises(AssertionError):
        mod([data["R0"].to(DEV), data["t0"].to(DEV)], lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"], lv["s0"], lv["s1"])


def test_bad_shapes_and_cpu_tensors_are_rejected():
    data = make_frame_pairs(2, 4, 16, 20, seed=6, n_levels=1)
    lv = {k: v.to(DEV) for k, v in data["levels"][0].items()}
    pose = (data["R0"].to(DEV), data["t0"].to(DEV))
    with pytest.raises(ValueError):
        A.uic_solve([dict(lv, s0=lv["s0"][:, :2])], pose)
    with pytest.raises(ValueError):
        A.uic_solve([dict(lv, K=lv["K"][:, :3])], pose)
    with pytest.raises(RuntimeError):
        A.uic_solve([dict(lv, x1=lv["x1"].cpu())], pose)
    mod = A.TrustRegionInverseWUncertainty(combine_icp=True)
    with pytest.raises(AssertionError):
        mod(list(pose), lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"], lv["s0"], lv["s1"])   # no depth given


def test_staged_kernel_with_depth_holes_and_out_of_view_motion():
    """The shape the staged-footprint kernel takes (C = 8, W % 4 == 0, W >= 60) with a third of the keyframe
    depth missing and a motion that pushes most lookups against the border: ring restarts, lanes outside
    the staged window and clamped footprints, all against the oracle with the mask bit-exact."""
    data = make_frame_pairs(3, 8, 36, 72, seed=18, n_levels=1)
    lv = data["levels"][0]
    hole = torch.rand(lv["invD0"].shape, generator=torch.Generator().manual_seed(2)) < 0.3
    lv["invD0"] = torch.where(hole, torch.zeros_like(lv["invD0"]), lv["invD0"]).contiguous()
    pose = _twist_to_pose(torch.tensor([[0.0, 0.35, 0.0, 0.4, 0.1, 0.0], [0.2, 0.0, 0.3, -0.3, 0.2, 0.1],
                                        [0.0, 0.0, 1.2, 0.0, 0.0, 0.0]]))      # the last one: 70 degrees in-plane
    res, trace = both(lv, pose, iters=2)
    assert trace[0]["occ"].float().mean() > 0.3
    check(res, trace, mask_slack=3)


def test_staged_and_plain_kernels_agree_and_misaligned_maps_fall_back():
    """staged_footprint only changes how the lookups are fetched: same masks, sums to rounding.  Maps that are
    not 16-byte aligned cannot be staged with 16-byte cp.async; the level then runs the plain kernel."""
    B, C, H, W = 4, 8, 60, 80
    data = make_frame_pairs(B, C, H, W, seed=44, n_levels=1)
    lv = levels_to(data["levels"], DEV)[0]
    pose = (data["R0"].to(DEV), data["t0"].to(DEV))
    a = A.uic_solve([lv], pose, iters=3, remove_tru_sigma=True, want_occ=True, staged_footprint=True)
    b = A.uic_solve([lv], pose, iters=3, remove_tru_sigma=True, want_occ=True, staged_footprint=False)
    # shift x1 / s1 by one float inside a larger allocation: still contiguous, no longer 16-byte aligned
    shifted = dict(lv)
    for k in ("x1", "s1"):
        buf = torch.empty(lv[k].numel() + 1, device=DEV)
        buf[1:] = lv[k].reshape(-1)
        shifted[k] = buf[1:].view_as(lv[k])
        assert shifted[k].data_ptr() % 16 != 0 and shifted[k].is_contiguous()
    c = A.uic_solve([shifted], pose, iters=3, remove_tru_sigma=True, want_occ=True, staged_footprint=True)
    torch.cuda.synchronize()
    assert torch.equal(a.occ[0], b.occ[0]) and torch.equal(c.occ[0], b.occ[0])
    assert (a.pose_hist - b.pose_hist).abs().max() < 1e-6
    assert torch.equal(c.sys_hist, b.sys_hist)          # the fallback IS the plain kernel
