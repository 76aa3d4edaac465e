"""The CPU oracle against fixtures produced by the reference itself (tests/golden/make_golden.py).

This is what pins the oracle: the reference ships no golden vectors of its own for this path.
"""
import pytest
import torch

from oracle import ic_oracle as O
from helpers import (TOL_POSE, ConvMEstimator, damping_mlp, frob_rel, level_inputs, load_golden)

UIC_CASES = ["uic_plain", "uic_trusigma", "uic_icp", "uic_masks", "uic_c8_wide"]


def run_uic(g, sampler, trace=None):
    f = g["flags"].tolist()
    lv = level_inputs(g)
    kw = {}
    if f[2]:
        kw = dict(obj_mask0=g["obj_mask0"].bool(), obj_mask1=g["obj_mask1"].bool())
    return O.uic_level((g["R0"], g["t0"]), lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"], lv["s0"],
                       lv["s1"], iters=f[3], remove_tru_sigma=bool(f[0]), combine_icp=bool(f[1]),
                       depth0=lv.get("depth0"), depth1=lv.get("depth1"), uncer_prop=True, sampler=sampler,
                       trace=trace, **kw), kw


@pytest.mark.parametrize("name", UIC_CASES)
@pytest.mark.parametrize("sampler", ["grid_sample", "explicit"])
def test_uic_level_matches_reference(name, sampler):
    g = load_golden(name)
    trace = []
    ((R, t), weights, A_last), kw = run_uic(g, sampler, trace)
    for i, rec in enumerate(trace):
        # validity masks: bit-exact, every iteration
        assert torch.equal(rec["occ"].to(torch.uint8), g["it_occ"][i])
        assert frob_rel(rec["A"], g["it_A"][i]) < 2e-6
        assert frob_rel(rec["b"], g["it_b"][i]) < 2e-5
    assert (R - g["R_out"]).abs().max() < 1e-6
    assert (t - g["t_out"]).abs().max() < 1e-6
    assert torch.allclose(weights, g["weights"])
    assert frob_rel(A_last, g["A_last"]) < 2e-6


@pytest.mark.parametrize("name", UIC_CASES)
def test_uic_residual_loss_matches_reference(name):
    g = load_golden(name)
    f = g["flags"].tolist()
    lv = level_inputs(g)
    kw = {}
    if f[2]:
        kw = dict(obj_mask0=g["obj_mask0"].bool(), obj_mask1=g["obj_mask1"].bool())
    loss = O.uic_residual_loss((g["R0"], g["t0"]), lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"],
                               lv["s0"], lv["s1"], remove_tru_sigma=bool(f[0]), combine_icp=bool(f[1]),
                               depth0=lv.get("depth0"), depth1=lv.get("depth1"), **kw)
    assert frob_rel(loss, g["res_loss"]) < 1e-6


@pytest.mark.parametrize("sampler", ["grid_sample", "explicit"])
def test_uic_pyramid_chain_matches_reference(sampler):
    g = load_golden("uic_pyramid")
    f = g["flags"].tolist()
    levels = [level_inputs(g, f"in{i}_") for i in range(4)]
    B = levels[0]["x0"].shape[0]
    trace = []
    pose, per_level = O.track_pyramid(levels, (torch.eye(3).repeat(B, 1, 1), torch.zeros(B, 3)), iters=f[3],
                                      remove_tru_sigma=bool(f[0]), sampler=sampler, trace=trace)
    for i in range(4):
        assert (per_level[i][0] - g[f"R_lvl{i}"]).abs().max() < 1e-6
        assert (per_level[i][1] - g[f"t_lvl{i}"]).abs().max() < 1e-6
        for j, rec in enumerate(trace[i]):
            assert torch.equal(rec["occ"].to(torch.uint8), g[f"it{i}_occ"][j])
    # the chain actually tracks: final error an order of magnitude below the initial one
    assert (pose[1] - g["t_gt"]).abs().max() < 0.2 * g["t_gt"].abs().max()


@pytest.mark.parametrize("name", ["uic_grad", "uic_grad_trusigma"])
@pytest.mark.parametrize("sampler", ["grid_sample", "explicit"])
def test_uic_autograd_matches_reference(name, sampler):
    g = load_golden(name)
    f = g["flags"].tolist()
    lv = level_inputs(g)
    leaves = {k: lv[k].clone().requires_grad_(True) for k in ("x0", "x1", "s0", "s1")}
    R0 = g["R0"].clone().requires_grad_(True)
    t0 = g["t0"].clone().requires_grad_(True)
    (R, t), _, A = O.uic_level((R0, t0), leaves["x0"], leaves["x1"], lv["invD0"], lv["invD1"], lv["K"],
                               leaves["s0"], leaves["s1"], iters=f[3], remove_tru_sigma=bool(f[0]),
                               uncer_prop=True, sampler=sampler)
    loss = (R * g["cR"]).sum() + (t * g["ct"]).sum() + (A * g["cA"]).sum()
    loss.backward()
    for k, v in leaves.items():
        assert frob_rel(v.grad, g["g_" + k]) < 1e-5, k
    assert frob_rel(R0.grad, g["g_R0"]) < 1e-5
    assert frob_rel(t0.grad, g["g_t0"]) < 1e-5


@pytest.mark.parametrize("name,solver", [("ic_plain", "Direct-Nodamping"), ("ic_resvol", "Direct-ResVol"),
                                         ("ic_deepic", "Direct-ResVol")])
@pytest.mark.parametrize("sampler", ["grid_sample", "explicit"])
def test_ic_level_matches_reference(name, solver, sampler):
    g = load_golden(name)
    f = g["flags"].tolist()
    lv = level_inputs(g)
    net = damping_mlp(g) if solver == "Direct-ResVol" else None
    mest = ConvMEstimator(g) if name == "ic_deepic" else None
    trace = []
    with torch.no_grad():
        (R, t), w = O.ic_level((g["R0"], g["t0"]), lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"],
                               iters=f[3], mest=mest, wPrior=g["wprior"], solver=solver, net=net,
                               sampler=sampler, trace=trace)
        loss = O.ic_residual_loss((g["R0"], g["t0"]), lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"],
                                  mest=mest, wPrior=g["wprior"], sampler=sampler)
    for j, rec in enumerate(trace):
        assert frob_rel(rec["H"], g["it_H"][j]) < 1e-5
        assert frob_rel(rec["b"], g["it_b"][j]) < 1e-4
    assert (R - g["R_out"]).abs().max() < TOL_POSE
    assert (t - g["t_out"]).abs().max() < TOL_POSE
    assert frob_rel(w, g["weights"]) < 1e-5
    assert frob_rel(loss, g["res_loss"]) < 1e-5


@pytest.mark.parametrize("sampler", ["grid_sample", "explicit"])
def test_ic_autograd_matches_reference(sampler):
    """Oracle autograd through a DeepIC level (conv M-estimator + residual-volume damping MLP) against the
    reference's own autograd."""
    g = load_golden("ic_grad")
    lv = level_inputs(g)
    net, mest = damping_mlp(g), ConvMEstimator(g)
    leaves = {k: lv[k].clone().requires_grad_(True) for k in ("x0", "x1")}
    R0 = g["R0"].clone().requires_grad_(True)
    t0 = g["t0"].clone().requires_grad_(True)
    (R, t), _ = O.ic_level((R0, t0), leaves["x0"], leaves["x1"], lv["invD0"], lv["invD1"], lv["K"],
                           iters=int(g["flags"][3]), mest=mest, wPrior=g["wprior"], solver="Direct-ResVol", net=net,
                           sampler=sampler)
    loss = (R * g["cR"]).sum() + (t * g["ct"]).sum()
    loss.backward()
    assert abs(loss.item() - g["loss"].item()) < 1e-5
    for k, v in leaves.items():
        assert frob_rel(v.grad, g["g_" + k]) < 1e-3, k
    assert frob_rel(R0.grad, g["g_R0"]) < 1e-3
    assert frob_rel(t0.grad, g["g_t0"]) < 1e-3
    assert frob_rel(mest.net[0].weight.grad, g["g_mest_conv0"]) < 1e-3
    assert frob_rel(net[0].weight.grad, g["g_solver_fc0"]) < 1e-3
    assert frob_rel(net[4].bias.grad, g["g_solver_fc2_bias"]) < 1e-3


def test_pose_epe_loss_matches_reference():
    """Oracle pose-pyramid loss (and its autograd) against the reference's compute_RT_EPE_loss, training call
    (60x80 resize, N poses) and evaluation call (full resolution, one pose)."""
    g = load_golden("pose_loss")
    B, _, H, W = g["depth"].shape
    rK = g["K"].clone() * torch.tensor([80.0 / W, 60.0 / H, 80.0 / W, 60.0 / H])
    R_est = g["R_est"].clone().requires_grad_(True)
    t_est = g["t_est"].clone().requires_grad_(True)
    loss = O.pose_epe_loss(R_est, t_est, g["R_gt"], g["t_gt"], g["rdepth"], rK, g["rinvalid"])
    (loss * g["w"]).sum().backward()
    assert frob_rel(loss, g["loss"]) < 1e-5 and loss[B - 1] == 0
    assert frob_rel(R_est.grad, g["g_R_est"]) < 1e-4 and frob_rel(t_est.grad, g["g_t_est"]) < 1e-4
    with torch.no_grad():
        ev = O.pose_epe_loss(g["R_est"][:, :1], g["t_est"][:, :1], g["R_gt"], g["t_gt"], g["depth"], g["K"],
                             g["invalid"].float())
    assert frob_rel(ev, g["loss_eval"]) < 1e-5


@pytest.mark.parametrize("kind", ["twoResidual", "MultiScale2w"])
def test_uic_level_with_a_learned_scaler_matches_the_installed_reference(kind):
    """The oracle's scale_func branch (alg:677-682) against the reference's own TrustRegionInverseWUncertainty with
    its ScaleNet, both on the CPU (baseline/_ref: the reference installed by recipe)."""
    from baseline import reference as REF
    if not REF.available():
        pytest.skip("baseline/_ref not installed")
    alg, _, _, _ = REF.modules()
    from deep_prob_feature_track_b200.synthetic import make_frame_pairs
    torch.manual_seed(5)
    net = alg.ScaleNet(kind, scale=0.05).eval()
    data = make_frame_pairs(2, 4, 48, 64, seed=13, n_levels=1, with_depth=True)
    lv = data["levels"][0]
    s0, s1 = lv["s0"].expand(-1, 4, -1, -1).contiguous(), lv["s1"].expand(-1, 4, -1, -1).contiguous()
    prior = torch.rand((2, 1, 24, 32), generator=torch.Generator().manual_seed(1))
    tr = alg.TrustRegionInverseWUncertainty(3, combine_icp=True, scale_func=net, remove_tru_sigma=True).eval()
    with torch.no_grad():
        (R_ref, t_ref), w_ref = tr([data["R0"], data["t0"]], lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"], s0, s1,
                                   wPrior=prior, depth0=lv["depth0"], depth1=lv["depth1"], vis_res=False)
        (R, t), w = O.uic_level((data["R0"], data["t0"]), lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"], s0, s1,
                                iters=3, remove_tru_sigma=True, combine_icp=True, depth0=lv["depth0"], depth1=lv["depth1"],
                                scale_func=net, wPrior=prior, sampler="grid_sample")
    assert frob_rel(w, w_ref) < 1e-5
    assert (R - R_ref).abs().max() < 1e-6 and (t - t_ref.reshape(-1, 3)).abs().max() < 1e-6
