"""Backward of the IC tracker (TrustRegionBase drop-in): each dpft_ic_*_backward entry point chained by
torch.autograd around the M-estimator CNN and the damping MLP, against the reference's own autograd
(fixture ic_grad) and against autograd through the CPU oracle."""
import pytest
import torch

from deep_prob_feature_track_b200 import algorithms as A
from helpers import TOL_GRAD, TOL_POSE, ConvMEstimator, damping_mlp, frob_rel, level_inputs, load_golden
from oracle import ic_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cuda_module(g, *, resvol, deep, iters):
    solver = A.DirectSolverNet("Direct-ResVol" if resvol else "Direct-Nodamping")
    if resvol:
        ref = damping_mlp(g)
        for i in range(3):
            solver.net[i][0].weight.data = ref[2 * i].weight.data.clone()
            solver.net[i][0].bias.data = ref[2 * i].bias.data.clone()
    mest = ConvMEstimator(g) if deep else None
    return A.TrustRegionBase(max_iter=iters, mEst_func=mest, solver_func=solver).to(DEV).eval()


def run_cuda(mod, g, lv, cR, ct, masks=(None, None)):
    leaves = {k: lv[k].to(DEV).requires_grad_(True) for k in ("x0", "x1")}
    R0 = g["R0"].to(DEV).requires_grad_(True)
    t0 = g["t0"].to(DEV).requires_grad_(True)
    m0, m1 = (m.to(DEV) if m is not None else None for m in masks)
    (R, t), _ = mod([R0, t0], leaves["x0"], leaves["x1"], lv["invD0"].to(DEV), lv["invD1"].to(DEV), lv["K"].to(DEV),
                    wPrior=g["wprior"].to(DEV), obj_mask0=m0, obj_mask1=m1)
    loss = (R * cR.to(DEV)).sum() + (t * ct.to(DEV)).sum()
    loss.backward()
    return R.detach().cpu(), t.detach().cpu(), {k: v.grad.cpu() for k, v in leaves.items()}, R0.grad.cpu(), t0.grad.cpu()


def test_deepic_gradients_match_reference():
    g = load_golden("ic_grad")
    lv = level_inputs(g)
    mod = cuda_module(g, resvol=True, deep=True, iters=int(g["flags"][3]))
    R, t, gm, gR0, gt0 = run_cuda(mod, g, lv, g["cR"], g["ct"])
    assert (R - g["R_out"]).abs().max() < TOL_POSE and (t - g["t_out"]).abs().max() < TOL_POSE
    for k in ("x0", "x1"):
        assert frob_rel(gm[k], g["g_" + k]) < TOL_GRAD, (k, frob_rel(gm[k], g["g_" + k]))
    assert frob_rel(gR0, g["g_R0"]) < TOL_GRAD and frob_rel(gt0, g["g_t0"]) < TOL_GRAD
    assert frob_rel(mod.mEstimator.net[0].weight.grad.cpu(), g["g_mest_conv0"]) < TOL_GRAD
    assert frob_rel(mod.directSolver.net[0][0].weight.grad.cpu(), g["g_solver_fc0"]) < TOL_GRAD
    assert frob_rel(mod.directSolver.net[2][0].bias.grad.cpu(), g["g_solver_fc2_bias"]) < TOL_GRAD


@pytest.mark.parametrize("name,resvol,deep,masks", [("ic_plain", False, False, False), ("ic_plain", False, False, True),
                                                    ("ic_resvol", True, False, False), ("ic_resvol", True, False, True),
                                                    ("ic_deepic", True, True, False)])
def test_gradients_match_oracle_autograd(name, resvol, deep, masks):
    g = load_golden(name)
    lv = level_inputs(g)
    B, C, H, W = lv["x0"].shape
    gen = torch.Generator().manual_seed(5)
    cR, ct = torch.randn((B, 3, 3), generator=gen), torch.randn((B, 3), generator=gen)
    m0 = m1 = None
    if masks:
        m0 = torch.rand((B, 1, H, W), generator=gen) > 0.2
        m1 = torch.rand((B, 1, H, W), generator=gen) > 0.2
    iters = 2
    mod = cuda_module(g, resvol=resvol, deep=deep, iters=iters)
    R, t, gm, gR0, gt0 = run_cuda(mod, g, lv, cR, ct, (m0, m1))
    # oracle
    net = damping_mlp(g) if resvol else None
    mest = ConvMEstimator(g) if deep else None
    leaves = {k: lv[k].clone().requires_grad_(True) for k in ("x0", "x1")}
    R0 = g["R0"].clone().requires_grad_(True)
    t0 = g["t0"].clone().requires_grad_(True)
    (Ro, to), _ = O.ic_level((R0, t0), leaves["x0"], leaves["x1"], lv["invD0"], lv["invD1"], lv["K"], iters=iters,
                             mest=mest, wPrior=g["wprior"], solver="Direct-ResVol" if resvol else "Direct-Nodamping",
                             net=net, obj_mask0=m0, obj_mask1=m1)
    ((Ro * cR).sum() + (to * ct).sum()).backward()
    assert (R - Ro.detach()).abs().max() < TOL_POSE and (t - to.detach()).abs().max() < TOL_POSE
    for k in ("x0", "x1"):
        assert frob_rel(gm[k], leaves[k].grad) < TOL_GRAD, (k, frob_rel(gm[k], leaves[k].grad))
    assert frob_rel(gR0, R0.grad) < TOL_GRAD and frob_rel(gt0, t0.grad) < TOL_GRAD
    if resvol:
        assert frob_rel(mod.directSolver.net[0][0].weight.grad.cpu(), net[0].weight.grad) < TOL_GRAD
    if deep:
        assert frob_rel(mod.mEstimator.net[0].weight.grad.cpu(), mest.net[0].weight.grad) < TOL_GRAD


def test_forward_honours_keyframe_mask_in_first_solve():
    """The first solve of a level uses the residual of the first warp, which includes obj_mask0 (alg:63-66, 80-82)."""
    g = load_golden("ic_plain")
    lv = level_inputs(g)
    B, C, H, W = lv["x0"].shape
    gen = torch.Generator().manual_seed(9)
    m0 = torch.rand((B, 1, H, W), generator=gen) > 0.3
    mod = cuda_module(g, resvol=False, deep=False, iters=3)
    with torch.no_grad():
        (R, t), _ = mod([g["R0"].to(DEV), g["t0"].to(DEV)], *(lv[k].to(DEV) for k in ("x0", "x1", "invD0", "invD1", "K")),
                        obj_mask0=m0.to(DEV))
        (Ro, to), _ = O.ic_level((g["R0"], g["t0"]), lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"], iters=3,
                                 obj_mask0=m0)
    assert (R.cpu() - Ro).abs().max() < TOL_POSE and (t.cpu() - to).abs().max() < TOL_POSE
