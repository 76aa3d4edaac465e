"""Context tensors of the learned networks inside the solver loop and what they feed:

  dpft_ic_context        the (B,4,H,W) input of DeepRobustEstimator('MultiScale2w') (alg:1471-1474) against the tensors
                         the reference concatenates; TrustRegionBase on that path against the reference fixture
  dpft_uic_icp_context   the two maps a learned ScaleNet reads (alg:677-680, 1535-1567) against the oracle
  icp_weight             a per-pixel scale of the point-to-plane term: against the scalar weight, against the oracle run
                         with the reference's own ScaleNet, and the whole tracker against the reference on CUDA
"""
import copy

import pytest
import torch
import torch.nn.functional as F

from baseline import reference as REF
from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import levels_to, make_frame_pairs
from helpers import TOL_POSE, TOL_SYS, ConvMEstimator, frob_rel, level_inputs, load_golden
from oracle import ic_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
needs_ref = pytest.mark.skipif(not REF.available(), reason="baseline/_ref not installed")


@pytest.fixture(autouse=True)
def exact_convolutions():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_ic_context_is_what_the_reference_concatenates():
    g = load_golden("ic_deepic")
    lv = {k: v.to(DEV) for k, v in level_inputs(g).items()}
    B, C, H, W = lv["x0"].shape
    assert C == 1
    lvl = A._IcLevel(lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"])
    rows = A.pack_pose((g["R0"], g["t0"])).to(DEV)
    wprior = g["wprior"].to(DEV)
    ctx = lvl.context(rows, wprior)
    r, _ = lvl.residual(rows, first=True)
    assert ctx.shape == (B, 4, H, W)
    assert torch.equal(ctx[:, 0:1], r.abs())
    assert torch.equal(ctx[:, 1:2], lv["x0"]) and torch.equal(ctx[:, 2:3], lv["x1"])
    up = F.interpolate(wprior, (H, W), mode="bilinear", align_corners=True)
    assert (ctx[:, 3:4] - up).abs().max() <= 1e-6 * up.abs().max()
    # and against the oracle's residual on the CPU
    px, py = O.pixel_rays(level_inputs(g)["K"], H, W)
    r_o, _ = O.ic_residual(g["R0"], g["t0"], *(level_inputs(g)[k] for k in ("invD0", "invD1", "x0", "x1")), px, py,
                           level_inputs(g)["K"])
    assert torch.equal(ctx[:, 0:1].cpu(), r_o.abs())


def test_ic_context_with_object_masks_and_no_prior():
    g = load_golden("ic_deepic")
    lv = {k: v.to(DEV) for k, v in level_inputs(g).items()}
    B, C, H, W = lv["x0"].shape
    gen = torch.Generator().manual_seed(4)
    m0 = (torch.rand((B, 1, H, W), generator=gen) > 0.2).to(DEV)
    m1 = (torch.rand((B, 1, H, W), generator=gen) > 0.2).to(DEV)
    lvl = A._IcLevel(lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"], m0, m1)
    rows = A.pack_pose((g["R0"], g["t0"])).to(DEV)
    ctx = lvl.context(rows, torch.ones((B, 1, 3, 4), device=DEV))
    r, _ = lvl.residual(rows, first=True)
    assert torch.equal(ctx[:, 0:1], r.abs())
    assert (ctx[:, 3] - 1).abs().max() <= 1e-6


def test_deepic_level_through_the_fused_context_matches_the_reference_fixture():
    """ic_deepic.npz was written by the reference's TrustRegionBase with its convolutional M-estimator; here the
    estimator's input comes from dpft_ic_context (D == 4 selects that path)."""
    g = load_golden("ic_deepic")
    lv = {k: v.to(DEV) for k, v in level_inputs(g).items()}
    mest = ConvMEstimator(g).to(DEV)
    mest.D = 4
    solver = A.DirectSolverNet("Direct-ResVol")
    from helpers import damping_mlp
    ref = damping_mlp(g)
    for i in range(3):
        solver.net[i][0].weight.data = ref[2 * i].weight.data.clone()
        solver.net[i][0].bias.data = ref[2 * i].bias.data.clone()
    mod = A.TrustRegionBase(max_iter=int(g["flags"][3]), mEst_func=mest, solver_func=solver).to(DEV).eval()
    pose = [g["R0"].to(DEV), g["t0"].to(DEV)]
    with torch.no_grad():
        assert mod._fused_context_ok(A._IcLevel(lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"]), g["wprior"], ())
        (R, t), w = mod(pose, lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"], wPrior=g["wprior"].to(DEV))
    assert (R.cpu() - g["R_out"]).abs().max() < TOL_POSE
    assert (t.cpu() - g["t_out"]).abs().max() < TOL_POSE
    assert frob_rel(w.cpu(), g["weights"]) < 1e-4


def _icp_level(seed=11, B=3, C=4, H=48, W=64):
    data = make_frame_pairs(B, C, H, W, seed=seed, n_levels=1, with_depth=True)
    return data, data["levels"][0]


@pytest.mark.parametrize("tru", [False, True])
def test_uic_icp_context_against_the_oracle(tru):
    data, lv = _icp_level()
    B, C, H, W = lv["x0"].shape
    R0, t0 = data["R0"], data["t0"]
    px, py = O.pixel_rays(lv["K"], H, W)
    s0 = lv["s0"].expand(-1, C, -1, -1)
    s1 = lv["s1"].expand(-1, C, -1, -1)
    wres, _, _, occ = O.uic_residuals(R0, t0, lv["invD0"], lv["invD1"], lv["x0"], lv["x1"], s0, s1, px, py, lv["K"],
                                      remove_tru_sigma=tru)
    V0, V1 = O.vertex_map(lv["depth0"], px, py), O.vertex_map(lv["depth1"], px, py)
    r_i, _, occ_i = O.icp_term(V0, V1, O.normal_map(V1), R0, t0, lv["K"])
    icp_r, feat_norm = A.uic_icp_context({k: v.to(DEV) for k, v in lv.items()}, (R0.to(DEV), t0.to(DEV)),
                                         remove_tru_sigma=tru)
    assert icp_r.shape == (B, 1, H, W) and feat_norm.shape == (B, 1, H, W)
    # masks: every masked pixel carries the reference's filler value, bit for bit
    assert torch.equal(icp_r.cpu() == O.EPS_UIC, r_i == O.EPS_UIC)
    assert occ_i.any() and (~occ_i).any()
    assert frob_rel(icp_r.cpu(), r_i) < 1e-5
    rtr = (wres * wres).sum(dim=1, keepdim=True)              # ScaleNet.compute_rtr (alg:1570-1574)
    assert frob_rel(feat_norm.cpu() ** 2, rtr) < 1e-5
    filler = (occ.expand(-1, 1, -1, -1))
    assert torch.allclose(feat_norm.cpu()[filler] ** 2, torch.full((), C * O.EPS_UIC ** 2), rtol=1e-5, atol=0)


def test_constant_weight_map_equals_the_scalar_weight():
    data, lv = _icp_level(seed=12)
    B, C, H, W = lv["x0"].shape
    dl = [{k: v.to(DEV) for k, v in lv.items()}]
    pose = (data["R0"].to(DEV), data["t0"].to(DEV))
    a = A.uic_solve(dl, pose, iters=3, combine_icp=True, w_icp=0.02)
    b = A.uic_solve(dl, pose, iters=3, combine_icp=True, icp_weight=[torch.full((B, 1, H, W), 0.02, device=DEV)])
    assert (a.pose[0] - b.pose[0]).abs().max() < 1e-6 and (a.pose[1] - b.pose[1]).abs().max() < 1e-6
    assert frob_rel(b.sys_hist, a.sys_hist) < 1e-5
    with pytest.raises(ValueError):
        A.uic_solve(dl, pose, iters=3, icp_weight=[torch.ones((B, 1, H, W), device=DEV)])      # needs combine_icp


@needs_ref
@pytest.mark.parametrize("kind", ["oneResidual", "twoResidual", "MultiScale2w", "expMultiScale"])
def test_learned_scaler_level_against_the_oracle(kind):
    """One level with the reference's own ScaleNet (random weights, eval mode): the oracle evaluates it on the full
    residual maps as the reference does, the module on the two maps the kernels write."""
    alg, _, _, _ = REF.modules()
    torch.manual_seed(5)
    net = alg.ScaleNet(kind, scale=0.05).eval()
    data, lv = _icp_level(seed=13, B=2)
    B, C, H, W = lv["x0"].shape
    prior = torch.rand((B, 1, H // 2, W // 2), generator=torch.Generator().manual_seed(1))
    s0 = lv["s0"].expand(-1, C, -1, -1)
    s1 = lv["s1"].expand(-1, C, -1, -1)
    trace = []
    with torch.no_grad():
        (R_o, t_o), w_o = O.uic_level((data["R0"], data["t0"]), lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"],
                                      s0, s1, iters=3, remove_tru_sigma=True, combine_icp=True, depth0=lv["depth0"],
                                      depth1=lv["depth1"], scale_func=net, wPrior=prior, trace=trace)
    mod = A.TrustRegionInverseWUncertainty(3, combine_icp=True, scale_func=copy.deepcopy(net).to(DEV),
                                           remove_tru_sigma=True, uncer_prop=True).to(DEV).eval()
    d = {k: v.to(DEV) for k, v in lv.items()}
    with torch.no_grad():
        (R, t), w, A_last = mod([data["R0"].to(DEV), data["t0"].to(DEV)], d["x0"], d["x1"], d["invD0"], d["invD1"],
                                d["K"], d["s0"].expand(-1, C, -1, -1), d["s1"].expand(-1, C, -1, -1),
                                wPrior=prior.to(DEV), depth0=d["depth0"], depth1=d["depth1"])
    assert w.shape == w_o.shape
    assert frob_rel(w.cpu(), w_o) < 1e-4
    assert frob_rel(A_last.cpu(), trace[-1]["A"]) < TOL_SYS
    assert (R.cpu() - R_o).abs().max() < TOL_POSE and (t.cpu() - t_o).abs().max() < TOL_POSE


@needs_ref
def test_tracker_with_a_learned_scaler_matches_the_reference_on_cuda():
    """LeastSquareTracking --combine_ICP --scaler MultiScale2w (the reference's learned ICP scale, LST:147-156):
    reference on CUDA against the patched copy."""
    flags = [f for f in REF.EVAL_TUM_FLAGS]
    flags[flags.index("--scaler") + 1] = "MultiScale2w"
    ref = REF.make_tracker(flags + ["--combine_ICP"], seed=0).to(DEV).eval()
    ours = A.patch_tracker(copy.deepcopy(ref)).eval()
    assert ours.tr_update0._learned_scaler()
    img0, img1, d0, d1, K = REF.synthetic_rgbd(2, 120, 160, seed=9, device=DEV)
    with torch.no_grad():
        R_ref, t_ref = ref(img0, img1, d0, d1, K)
        R, t = ours(img0, img1, d0, d1, K)
    assert t_ref.abs().max() > 1e-5
    assert (R - R_ref).abs().max() < 1e-5, (R - R_ref).abs().max()
    assert (t - t_ref).abs().max() < 1e-5, (t - t_ref).abs().max()
