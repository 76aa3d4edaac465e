"""CUDA U_IC forward vs the CPU oracle and vs the reference-generated golden fixtures.

Bars (BASELINE.json north_star / SURVEY.md 8c): validity masks bit-exact, J^T W J and J^T W r within
1e-4 Frobenius-relative, poses within 1e-5.  Everything goes through the C ABI (uic_solve -> ctypes).
"""
import pytest
import torch

from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import levels_to, make_frame_pairs
from helpers import TOL_POSE, TOL_SYS, frob_rel, level_inputs, load_golden
from oracle import ic_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FUSED = True
SINGLE = True
STAGED = False


@pytest.fixture(autouse=True, params=[(True, True, False), (True, False, False), (False, False, False), (True, False, True)],
                ids=["single-launch", "launch-per-iteration", "materialised-gradients", "staged-footprint"])
def forward_path(request, monkeypatch):
    """Every test runs on the four forward paths: the single cooperative launch, the same fused kernel
    launched once per iteration, the variant with gradients materialised once per level, and the fused kernel
    with the lookup footprint staged in shared memory (levels that qualify: C == 8, W % 4 == 0, W >= 48)."""
    import sys
    monkeypatch.setattr(sys.modules[__name__], "FUSED", request.param[0])
    monkeypatch.setattr(sys.modules[__name__], "SINGLE", request.param[1])
    monkeypatch.setattr(sys.modules[__name__], "STAGED", request.param[2])


def perturbed(B, seed, scale=0.01):
    from deep_prob_feature_track_b200.synthetic import _twist_to_pose
    g = torch.Generator().manual_seed(seed)
    return _twist_to_pose((torch.rand((B, 6), generator=g) * 2 - 1) * scale)


def run_cuda(levels, pose, **kw):
    kw.setdefault("fused_sobel", FUSED)
    kw.setdefault("single_launch", SINGLE)
    kw.setdefault("staged_footprint", STAGED)
    kw.setdefault("queue", False)      # the work-queue path has its own file (test_uic_queue_gpu.py)
    res = A.uic_solve(levels_to(levels, DEV), (pose[0].to(DEV), pose[1].to(DEV)), want_occ=True, **kw)
    torch.cuda.synchronize()
    return res


def compare_level(res, trace, k0, lvl, iters, *, exact_pose_inputs):
    """Compare iterations k0..k0+iters of a CUDA run with the oracle's trace of one level."""
    flips = 0
    for it, rec in enumerate(trace):
        Ac, bc = A.unpack_system(res.sys_hist[k0 + it].cpu())
        assert frob_rel(Ac, rec["A"]) < TOL_SYS, (lvl, it, frob_rel(Ac, rec["A"]))
        assert frob_rel(bc, rec["b"]) < TOL_SYS, (lvl, it, frob_rel(bc, rec["b"]))
        occ_c = res.occ[lvl][it].cpu()
        occ_o = rec["occ"][:, 0].to(torch.uint8)
        n = int((occ_c != occ_o).sum())
        if it == 0 and exact_pose_inputs:
            assert n == 0, f"mask not bit-exact at level {lvl} it 0: {n} flips"
        flips += n
        Rc, tc = A.unpack_pose(res.pose_hist[k0 + it].cpu())
        assert (Rc - rec["R"]).abs().max() < TOL_POSE
        assert (tc - rec["t"]).abs().max() < TOL_POSE
    return flips


@pytest.mark.parametrize("name", ["uic_plain", "uic_trusigma", "uic_masks", "uic_c8_wide", "uic_icp"])
def test_golden_single_level(name):
    """Inputs and expected outputs come from the reference itself (tests/golden/make_golden.py)."""
    g = load_golden(name)
    f = g["flags"].tolist()
    lv = level_inputs(g)
    kw = {}
    okw = {}
    if f[2]:
        kw = dict(obj_mask0=[g["obj_mask0"].to(DEV)], obj_mask1=[g["obj_mask1"].to(DEV)])
        okw = dict(obj_mask0=g["obj_mask0"].bool(), obj_mask1=g["obj_mask1"].bool())
    res = run_cuda([lv], (g["R0"], g["t0"]), iters=f[3], remove_tru_sigma=bool(f[0]), combine_icp=bool(f[1]), **kw)
    assert int(res.status.item()) == 0
    # forward_residuals at the starting pose (reference alg:725-786)
    loss = A.uic_residual_loss({k: v.to(DEV) for k, v in lv.items()}, (g["R0"].to(DEV), g["t0"].to(DEV)),
                               remove_tru_sigma=bool(f[0]), combine_icp=bool(f[1]),
                               obj_mask0=kw["obj_mask0"][0] if kw else None, obj_mask1=kw["obj_mask1"][0] if kw else None)
    assert frob_rel(loss.cpu(), g["res_loss"]) < 1e-4, frob_rel(loss.cpu(), g["res_loss"])
    # against the reference's recorded iterations
    total_flips = 0
    for it in range(f[3]):
        Ac, bc = A.unpack_system(res.sys_hist[it].cpu())
        assert frob_rel(Ac, g["it_A"][it]) < TOL_SYS
        assert frob_rel(bc, g["it_b"][it]) < TOL_SYS
        total_flips += int((res.occ[0][it].cpu() != g["it_occ"][it][:, 0]).sum())
    assert total_flips == 0, f"{total_flips} mask flips against the reference"
    R, t = (x.cpu() for x in res.pose)
    assert (R - g["R_out"]).abs().max() < TOL_POSE
    assert (t - g["t_out"]).abs().max() < TOL_POSE
    # and against the oracle, bit-exact masks
    trace = []
    O.uic_level((g["R0"], g["t0"]), lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"], lv["s0"], lv["s1"],
                iters=f[3], remove_tru_sigma=bool(f[0]), combine_icp=bool(f[1]), depth0=lv.get("depth0"),
                depth1=lv.get("depth1"), trace=trace, **okw)
    assert compare_level(res, trace, 0, 0, f[3], exact_pose_inputs=True) == 0


def test_golden_pyramid_chain():
    g = load_golden("uic_pyramid")
    f = g["flags"].tolist()
    levels = [level_inputs(g, f"in{i}_") for i in range(4)]
    B = levels[0]["x0"].shape[0]
    res = run_cuda(levels, (torch.eye(3).repeat(B, 1, 1), torch.zeros(B, 3)), iters=f[3],
                   remove_tru_sigma=bool(f[0]))
    assert int(res.status.item()) == 0
    flips = 0
    for i in range(4):
        R, t = (x.cpu() for x in res.level_pose(i))
        assert (R - g[f"R_lvl{i}"]).abs().max() < TOL_POSE
        assert (t - g[f"t_lvl{i}"]).abs().max() < TOL_POSE
        for it in range(f[3]):
            Ac, bc = A.unpack_system(res.sys_hist[i * f[3] + it].cpu())
            assert frob_rel(Ac, g[f"it{i}_A"][it]) < TOL_SYS
            assert frob_rel(bc, g[f"it{i}_b"][it]) < TOL_SYS
            flips += int((res.occ[i][it].cpu() != g[f"it{i}_occ"][it][:, 0]).sum())
    # later iterations start from poses that differ in the last bits, so allow threshold-adjacent flips
    assert flips <= 2, flips


SHAPES = [
    # B, C, H, W, remove_tru_sigma
    (3, 8, 60, 80, True),
    (2, 8, 23, 37, False),     # ragged: not multiples of the tile
    (1, 1, 30, 40, True),      # single channel (RGB / DeepIC feature shape)
    (2, 3, 16, 31, True),      # odd channel count -> 1-channel chunks
    (2, 16, 24, 32, True),     # two 8-channel chunks
    (5, 4, 15, 20, False),     # coarsest TUM level
    (2, 2, 9, 61, True),       # wider than two warp tiles, fewer rows than a tile
]


@pytest.mark.parametrize("B,C,H,W,tru", SHAPES)
def test_single_level_vs_oracle(B, C, H, W, tru):
    data = make_frame_pairs(B, C, H, W, seed=100 + B * C + H, n_levels=1)
    lv = data["levels"][0]
    pose = perturbed(B, 7 + H)
    res = run_cuda([lv], pose, iters=3, remove_tru_sigma=tru)
    assert int(res.status.item()) == 0
    trace = []
    (R, t), _ = O.uic_level(pose, lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"], lv["s0"], lv["s1"],
                            iters=3, remove_tru_sigma=tru, trace=trace)
    flips = compare_level(res, trace, 0, 0, 3, exact_pose_inputs=True)
    assert flips <= max(1, B * H * W // 5000), flips
    Rc, tc = (x.cpu() for x in res.pose)
    assert (Rc - R).abs().max() < TOL_POSE and (tc - t).abs().max() < TOL_POSE


def test_masks_bit_exact_over_many_poses():
    """One iteration from the same pose on both sides: the integer mask must match bit for bit."""
    B, C, H, W = 8, 4, 48, 64
    data = make_frame_pairs(B, C, H, W, seed=5, n_levels=1)
    lv = data["levels"][0]
    for s, scale in ((1, 0.002), (2, 0.02), (3, 0.08), (4, 0.3)):
        pose = perturbed(B, s, scale)
        res = run_cuda([lv], pose, iters=1, remove_tru_sigma=True)
        trace = []
        O.uic_level(pose, lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"], lv["s0"], lv["s1"], iters=1,
                    remove_tru_sigma=True, trace=trace)
        assert torch.equal(res.occ[0][0].cpu(), trace[0]["occ"][:, 0].to(torch.uint8)), scale


@pytest.mark.parametrize("scale", [0.005, 0.05, 0.15, 0.4])
def test_eight_channel_level_under_small_and_large_motion(scale):
    """C = 8, W % 4 == 0: the shape the staged-footprint kernel takes.  Large motions spread the lookups of a
    warp row over more source rows / columns than its shared-memory ring holds and push them against the image
    border, so the per-lane direct loads and the ring restarts are exercised; results must not depend on it."""
    B, C, H, W = 4, 8, 44, 64
    data = make_frame_pairs(B, C, H, W, seed=31, n_levels=1)
    lv = data["levels"][0]
    pose = perturbed(B, 11, scale)
    res = run_cuda([lv], pose, iters=2, remove_tru_sigma=True)
    trace = []
    O.uic_level(pose, lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"], lv["s0"], lv["s1"], iters=2,
                remove_tru_sigma=True, trace=trace)
    assert torch.equal(res.occ[0][0].cpu(), trace[0]["occ"][:, 0].to(torch.uint8))
    for it in range(2):
        Ac, bc = A.unpack_system(res.sys_hist[it].cpu())
        assert frob_rel(Ac, trace[it]["A"]) < TOL_SYS and frob_rel(bc, trace[it]["b"]) < TOL_SYS, (scale, it)


def test_saturated_sigma_masks():
    """sigma clamped like the reference's laplacian head (exp(clamp(.,-3,3))): whole regions sit on the
    batch extremes and must be masked exactly as the oracle masks them."""
    B, C, H, W = 3, 4, 40, 52
    data = make_frame_pairs(B, C, H, W, seed=77, n_levels=1)
    lv = data["levels"][0]
    for k in ("s0", "s1"):
        lv[k] = lv[k].clamp(0.8, 1.25).contiguous()
    pose = perturbed(B, 3)
    res = run_cuda([lv], pose, iters=2, remove_tru_sigma=True)
    trace = []
    O.uic_level(pose, lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"], lv["s0"], lv["s1"], iters=2,
                remove_tru_sigma=True, trace=trace)
    assert trace[0]["occ"].float().mean() > 0.15          # the case really exercises the mask
    assert compare_level(res, trace, 0, 0, 2, exact_pose_inputs=True) <= 1


def test_full_size_batch_properties():
    """BASELINE config 2 size (B=64, C=8, 120x160, 4 levels): determinism, batch-order equivariance and,
    without the batch-global sigma test, independence of every pair from its batch-mates."""
    B, C, H, W = 64, 8, 120, 160
    data = make_frame_pairs(B, C, H, W, seed=1234, n_levels=4)
    levels = levels_to(data["levels"], DEV)
    pose = (data["R0"].to(DEV), data["t0"].to(DEV))
    r1 = A.uic_solve(levels, pose, iters=3, remove_tru_sigma=True, fused_sobel=FUSED, single_launch=SINGLE, staged_footprint=STAGED, queue=False)
    r2 = A.uic_solve(levels, pose, iters=3, remove_tru_sigma=True, fused_sobel=FUSED, single_launch=SINGLE, staged_footprint=STAGED, queue=False)
    torch.cuda.synchronize()
    assert int(r1.status.item()) == 0
    assert torch.equal(r1.pose_hist, r2.pose_hist) and torch.equal(r1.sys_hist, r2.sys_hist)
    # converges towards the motion the data was made with
    R, t = r1.pose
    err0 = data["t_gt"].abs().max().item()
    assert (t.cpu() - data["t_gt"]).abs().max().item() < 0.25 * err0
    # permuting the batch permutes the result (bitwise)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(0)).to(DEV)
    lv_p = [{k: v[perm].contiguous() for k, v in lv.items()} for lv in levels]
    r3 = A.uic_solve(lv_p, (pose[0][perm], pose[1][perm]), iters=3, remove_tru_sigma=True, fused_sobel=FUSED, single_launch=SINGLE, staged_footprint=STAGED, queue=False)
    if STAGED:
        # the balanced tiling gives some pairs one CTA more than others depending on their position in the batch, so
        # a permuted pair is summed in a different order: equal to rounding instead of bitwise
        assert (r3.pose_hist - r1.pose_hist[:, perm]).abs().max() < 1e-6
    else:
        assert torch.equal(r3.pose_hist, r1.pose_hist[:, perm])
    # no batch coupling without remove_tru_sigma: swapping the batch-mates of the first five pairs for
    # other data leaves their rows bitwise unchanged, and a smaller batch (different tiling, so a
    # different summation order) agrees to rounding
    r4 = A.uic_solve(levels, pose, iters=3, remove_tru_sigma=False, fused_sobel=FUSED, single_launch=SINGLE, staged_footprint=STAGED, queue=False)
    mixed = [{k: torch.cat((v[:5], v[5:].flip(0))).contiguous() for k, v in lv.items()} for lv in levels]
    r5 = A.uic_solve(mixed, pose, iters=3, remove_tru_sigma=False, fused_sobel=FUSED, single_launch=SINGLE, staged_footprint=STAGED, queue=False)
    assert torch.equal(r5.pose_hist[:, :5], r4.pose_hist[:, :5])
    sub = [{k: v[:5].contiguous() for k, v in lv.items()} for lv in levels]
    r7 = A.uic_solve(sub, (pose[0][:5], pose[1][:5]), iters=3, remove_tru_sigma=False, fused_sobel=FUSED, single_launch=SINGLE, staged_footprint=STAGED, queue=False)
    assert (r7.pose_hist - r4.pose_hist[:, :5]).abs().max() < 1e-6
    # PDL on/off is only a scheduling difference
    r6 = A.uic_solve(levels, pose, iters=3, remove_tru_sigma=True, pdl=False, fused_sobel=FUSED, single_launch=SINGLE, staged_footprint=STAGED, queue=False)
    assert torch.equal(r6.pose_hist, r1.pose_hist)


def test_full_size_vs_oracle():
    """The whole B=64 120x160 4-level solve against the oracle (a few seconds of CPU)."""
    B, C, H, W = 64, 8, 120, 160
    data = make_frame_pairs(B, C, H, W, seed=4321, n_levels=4)
    res = run_cuda(data["levels"], (data["R0"], data["t0"]), iters=3, remove_tru_sigma=True, fused_sobel=FUSED, single_launch=SINGLE, staged_footprint=STAGED, queue=False)
    trace = []
    with torch.no_grad():
        pose, per_level = O.track_pyramid(data["levels"], (data["R0"], data["t0"]), iters=3,
                                          remove_tru_sigma=True, trace=trace, reduction="einsum")
    flips = 0
    for i in range(4):
        flips += compare_level(res, trace[i], 3 * i, i, 3, exact_pose_inputs=(i == 0))
    total = sum(3 * B * lv["x0"].shape[2] * lv["x0"].shape[3] for lv in data["levels"])
    assert flips <= total * 1e-5, (flips, total)
    R, t = (x.cpu() for x in res.pose)
    assert (R - pose[0]).abs().max() < TOL_POSE and (t - pose[1]).abs().max() < TOL_POSE
    # the generator repeats ONE uncertainty map to the C channels, like the reference's encoder: handing that map over
    # un-repeated (DPFT_SIGMA_BROADCAST) must give the same solve at the full size and with the balanced tiling
    one = [dict(lv, s0=lv["s0"][:, :1].contiguous(), s1=lv["s1"][:, :1].contiguous()) for lv in data["levels"]]
    assert all(torch.equal(lv["s0"], o["s0"].expand_as(lv["s0"])) for lv, o in zip(data["levels"], one))
    res1 = run_cuda(one, (data["R0"], data["t0"]), iters=3, remove_tru_sigma=True)
    assert torch.equal(res1.occ[0][0], res.occ[0][0])            # same starting pose: bit for bit
    for i in range(1, 4):                                        # later levels start from poses equal to rounding
        assert int((res1.occ[i] != res.occ[i]).sum()) <= 5e-5 * res.occ[i].numel(), i
    assert (res1.pose_hist - res.pose_hist).abs().max() < 1e-6
    assert frob_rel(res1.sys_hist.cpu(), res.sys_hist.cpu()) < 1e-5


def test_module_surface_matches_reference_signature():
    B, C, H, W = 2, 4, 24, 32
    g = load_golden("uic_trusigma")
    lv = {k: v.to(DEV) for k, v in level_inputs(g).items()}
    mod = A.TrustRegionInverseWUncertainty(max_iter=3, remove_tru_sigma=True, uncer_prop=True)
    with torch.no_grad():
        pose, weights, JtWJ = mod([g["R0"].to(DEV), g["t0"].to(DEV).view(B, 3, 1)], lv["x0"], lv["x1"], lv["invD0"],
                                  lv["invD1"], lv["K"], lv["s0"], lv["s1"], wPrior=None, vis_res=False)
    assert pose[0].shape == (B, 3, 3) and pose[1].shape == (B, 3)
    assert weights.shape == (B, C, H, W) and bool((weights == 1).all())
    assert frob_rel(JtWJ.cpu(), g["A_last"]) < TOL_SYS
    assert (pose[1].cpu() - g["t_out"]).abs().max() < TOL_POSE
    for name in ("max_iterations", "mEstimator", "directSolver", "timers", "scale_func", "combine_icp",
                 "uncer_prop", "remove_tru_sigma"):
        assert hasattr(mod, name)
    with pytest.raises(RuntimeError):
        mod([g["R0"], g["t0"]], *(level_inputs(g)[k] for k in ("x0", "x1", "invD0", "invD1", "K", "s0", "s1")))


@pytest.mark.parametrize("shape", [(3, 8, 60, 80), (2, 8, 120, 160), (2, 4, 24, 32), (2, 3, 16, 31)])
def test_single_uncertainty_map_equals_the_repeated_tensor(shape):
    """sigma given as (B,1,h,w) -- what the reference's encoder emits before `repeat` (alg:1425-1427) -- against
    the same map repeated to C channels and against the oracle: masks identical, sums to rounding."""
    B, C, H, W = shape
    data = make_frame_pairs(B, C, H, W, seed=61, n_levels=1)
    lv = dict(data["levels"][0])
    lv["s0"] = lv["s0"][:, :1].repeat(1, C, 1, 1).contiguous()
    lv["s1"] = lv["s1"][:, :1].repeat(1, C, 1, 1).contiguous()
    one = dict(lv, s0=lv["s0"][:, :1].contiguous(), s1=lv["s1"][:, :1].contiguous())
    pose = perturbed(B, 5, 0.02)
    full = run_cuda([lv], pose, iters=3, remove_tru_sigma=True)
    bc = run_cuda([one], pose, iters=3, remove_tru_sigma=True)
    assert int(bc.status.item()) == 0
    assert torch.equal(bc.occ[0], full.occ[0])
    assert frob_rel(bc.sys_hist.cpu(), full.sys_hist.cpu()) < 1e-5
    assert (bc.pose_hist - full.pose_hist).abs().max() < 1e-6
    trace = []
    O.uic_level(pose, lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"], lv["s0"], lv["s1"], iters=3,
                remove_tru_sigma=True, trace=trace)
    compare_level(bc, trace, 0, 0, 3, exact_pose_inputs=True)


def test_icp_variant_is_bitwise_reproducible():
    """The point-to-plane sums are folded in CTA order by each pair's last CTA (no float atomics on the solver's path):
    repeated solves give the same floats."""
    from deep_prob_feature_track_b200.synthetic import make_frame_pairs
    d = make_frame_pairs(8, 8, 120, 160, seed=31, n_levels=4, with_depth=True)
    levels = [{k: v.to(DEV) for k, v in lv.items()} for lv in d["levels"]]
    pose = (d["R0"].to(DEV), d["t0"].to(DEV))
    runs = [A.uic_solve(levels, pose, iters=3, remove_tru_sigma=True, combine_icp=True, w_icp=0.01) for _ in range(6)]
    for r in runs[1:]:
        assert torch.equal(r.pose_hist, runs[0].pose_hist) and torch.equal(r.sys_hist, runs[0].sys_hist)
    runs[0].raise_if_bad()
