"""CUDA U_IC backward (dpft_uic_backward through torch.autograd) against reference autograd fixtures and
against autograd on the CPU oracle.  Tolerance: 1e-3 Frobenius-relative on every gradient (the chain goes
through 2-6 unrolled solves; fp32 reference-vs-oracle noise is 1e-6)."""
import pytest
import torch

from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import _twist_to_pose, make_frame_pairs
from helpers import TOL_GRAD, frob_rel, level_inputs, load_golden
from oracle import ic_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cuda_grads(levels, R0, t0, iters, tru, loss_fn, icp=False, w_icp=0.01):
    lv_dev = []
    leaves = []
    for lv in levels:
        d = {k: v.to(DEV) for k, v in lv.items()}
        for k in ("x0", "x1", "s0", "s1"):
            d[k] = d[k].clone().requires_grad_(True)
        lv_dev.append(d)
        leaves.append({k: d[k] for k in ("x0", "x1", "s0", "s1")})
    R = R0.to(DEV).clone().requires_grad_(True)
    t = t0.to(DEV).clone().requires_grad_(True)
    outs = A.uic_track(lv_dev, (R, t), iters=iters, remove_tru_sigma=tru, combine_icp=icp, w_icp=w_icp)
    loss = loss_fn(outs)
    loss.backward()
    torch.cuda.synchronize()
    return loss.item(), leaves, R.grad.cpu(), t.grad.cpu(), outs


@pytest.mark.parametrize("name", ["uic_grad", "uic_grad_trusigma"])
def test_against_reference_autograd(name):
    """Fixture made by running the reference's own autograd (tests/golden/make_golden.py: uic_gradients)."""
    g = load_golden(name)
    f = g["flags"].tolist()
    lv = level_inputs(g)
    cR, ct, cA = g["cR"].to(DEV), g["ct"].to(DEV), g["cA"].to(DEV)

    def loss_fn(outs):
        R, t, Am = outs[0]
        return (R * cR).sum() + (t * ct).sum() + (Am * cA).sum()

    loss, leaves, gR, gt, _ = cuda_grads([lv], g["R0"], g["t0"], f[3], bool(f[0]), loss_fn)
    assert abs(loss - g["loss"].item()) < 1e-4 * max(1.0, abs(g["loss"].item()))
    for k, v in leaves[0].items():
        assert frob_rel(v.grad.cpu(), g["g_" + k]) < TOL_GRAD, (k, frob_rel(v.grad.cpu(), g["g_" + k]))
    assert frob_rel(gR, g["g_R0"]) < TOL_GRAD
    assert frob_rel(gt, g["g_t0"]) < TOL_GRAD


@pytest.mark.parametrize("B,C,H,W,n_levels,tru", [(2, 8, 24, 32, 1, True), (2, 4, 40, 56, 3, True),
                                                  (3, 3, 17, 23, 2, False), (1, 1, 20, 28, 1, True)])
def test_against_oracle_autograd(B, C, H, W, n_levels, tru):
    """Loss = the per-level pose functional train.py uses in spirit: every level's pose is an output."""
    data = make_frame_pairs(B, C, H, W, seed=50 + C + n_levels, n_levels=n_levels)
    gen = torch.Generator().manual_seed(9)
    R0, t0 = _twist_to_pose((torch.rand((B, 6), generator=gen) * 2 - 1) * 0.01)
    cs = [(torch.randn((B, 3, 3), generator=gen), torch.randn((B, 3), generator=gen)) for _ in range(n_levels)]

    def loss_cuda(outs):
        return sum((R * c[0].to(DEV)).sum() + (t * c[1].to(DEV)).sum() for (R, t, _), c in zip(outs, cs))

    loss, leaves, gR, gt, _ = cuda_grads(data["levels"], R0, t0, 3, tru, loss_cuda)

    o_levels = []
    for lv in data["levels"]:
        d = dict(lv)
        for k in ("x0", "x1", "s0", "s1"):
            d[k] = lv[k].clone().requires_grad_(True)
        o_levels.append(d)
    Ro, to = R0.clone().requires_grad_(True), t0.clone().requires_grad_(True)
    _, per_level = O.track_pyramid(o_levels, (Ro, to), iters=3, remove_tru_sigma=tru, reduction="einsum")
    loss_o = sum((R * c[0]).sum() + (t * c[1]).sum() for (R, t), c in zip(per_level, cs))
    loss_o.backward()
    assert abs(loss - loss_o.item()) < 1e-4 * max(1.0, abs(loss_o.item()))
    for l in range(n_levels):
        for k in ("x0", "x1", "s0", "s1"):
            e = frob_rel(leaves[l][k].grad.cpu(), o_levels[l][k].grad)
            assert e < TOL_GRAD, (l, k, e)
    assert frob_rel(gR, Ro.grad) < TOL_GRAD
    assert frob_rel(gt, to.grad) < TOL_GRAD


def test_module_trains_like_the_reference_module():
    """nn.Module surface in train mode: gradients reach the feature maps and the incoming pose."""
    g = load_golden("uic_grad")
    f = g["flags"].tolist()
    lv = {k: v.to(DEV) for k, v in level_inputs(g).items()}
    for k in ("x0", "x1", "s0", "s1"):
        lv[k].requires_grad_(True)
    R0 = g["R0"].to(DEV).requires_grad_(True)
    t0 = g["t0"].to(DEV).view(-1, 3, 1).requires_grad_(True)
    mod = A.TrustRegionInverseWUncertainty(max_iter=f[3], uncer_prop=True).train()
    (R, t), w, Am = mod([R0, t0], lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"], lv["s0"], lv["s1"])
    loss = (R * g["cR"].to(DEV)).sum() + (t * g["ct"].to(DEV)).sum() + (Am * g["cA"].to(DEV)).sum()
    loss.backward()
    assert frob_rel(lv["x1"].grad.cpu(), g["g_x1"]) < TOL_GRAD
    assert frob_rel(t0.grad.cpu().view(-1, 3), g["g_t0"]) < TOL_GRAD


@pytest.mark.parametrize("n_levels,w_icp", [(1, 0.01), (2, 0.01), (1, 1.0)])
def test_icp_term_backward_against_oracle_autograd(n_levels, w_icp):
    """combine_icp in training (train_tum_feature_icp.sh): the point-to-plane term adds a pose-gradient path."""
    B, C, H, W = 2, 4, 32, 44
    data = make_frame_pairs(B, C, H, W, seed=60 + n_levels, n_levels=n_levels, with_depth=True)
    gen = torch.Generator().manual_seed(3)
    R0, t0 = _twist_to_pose((torch.rand((B, 6), generator=gen) * 2 - 1) * 0.01)
    cs = [(torch.randn((B, 3, 3), generator=gen), torch.randn((B, 3), generator=gen)) for _ in range(n_levels)]

    def loss_cuda(outs):
        return sum((R * c[0].to(DEV)).sum() + (t * c[1].to(DEV)).sum() for (R, t, _), c in zip(outs, cs))

    loss, leaves, gR, gt, _ = cuda_grads(data["levels"], R0, t0, 2, True, loss_cuda, icp=True, w_icp=w_icp)
    o_levels = []
    for lv in data["levels"]:
        d = dict(lv)
        for k in ("x0", "x1", "s0", "s1"):
            d[k] = lv[k].clone().requires_grad_(True)
        o_levels.append(d)
    Ro, to = R0.clone().requires_grad_(True), t0.clone().requires_grad_(True)
    # w_icp = 1 makes the point-to-plane term dominate the system, so its gradient path is really exercised
    _, per_level = O.track_pyramid(o_levels, (Ro, to), iters=2, remove_tru_sigma=True, combine_icp=True,
                                   scale_func=lambda r, w, prior: torch.ones_like(r) * w_icp, reduction="einsum")
    loss_o = sum((R * c[0]).sum() + (t * c[1]).sum() for (R, t), c in zip(per_level, cs))
    loss_o.backward()
    assert abs(loss - loss_o.item()) < 1e-4 * max(1.0, abs(loss_o.item()))
    for l in range(n_levels):
        for k in ("x0", "x1", "s0", "s1"):
            e = frob_rel(leaves[l][k].grad.cpu(), o_levels[l][k].grad)
            assert e < TOL_GRAD, (l, k, e)
    assert frob_rel(gR, Ro.grad) < TOL_GRAD, frob_rel(gR, Ro.grad)
    assert frob_rel(gt, to.grad) < TOL_GRAD, frob_rel(gt, to.grad)


def _full_size_case(B=4, seed=91):
    data = make_frame_pairs(B, 8, 120, 160, seed=seed, n_levels=4)
    gen = torch.Generator().manual_seed(9)
    R0, t0 = _twist_to_pose((torch.rand((B, 6), generator=gen) * 2 - 1) * 0.01)
    cs = [(torch.randn((B, 3, 3), generator=gen), torch.randn((B, 3), generator=gen)) for _ in range(4)]
    return data, R0, t0, cs


def test_full_resolution_pyramid_against_oracle_autograd():
    """The training configuration of BASELINE.json (120x160, 8 channels, 4 levels x 3 iterations, remove_tru_sigma) at
    B = 4: every gradient against autograd through the CPU oracle.  Measured 2e-6 (pose) .. 1.5e-4 (sigma0 of the 30x40
    level; fp32 noise through twelve unrolled solves, the coarse levels see all of them); gate at 5e-4, half of TOL_GRAD."""
    data, R0, t0, cs = _full_size_case()

    def loss_cuda(outs):
        return sum((R * c[0].to(DEV)).sum() + (t * c[1].to(DEV)).sum() for (R, t, _), c in zip(outs, cs))

    loss, leaves, gR, gt, _ = cuda_grads(data["levels"], R0, t0, 3, True, loss_cuda)
    o_levels = []
    for lv in data["levels"]:
        d = dict(lv)
        for k in ("x0", "x1", "s0", "s1"):
            d[k] = lv[k].clone().requires_grad_(True)
        o_levels.append(d)
    Ro, to = R0.clone().requires_grad_(True), t0.clone().requires_grad_(True)
    _, per_level = O.track_pyramid(o_levels, (Ro, to), iters=3, remove_tru_sigma=True, reduction="einsum")
    loss_o = sum((R * c[0]).sum() + (t * c[1]).sum() for (R, t), c in zip(per_level, cs))
    loss_o.backward()
    assert abs(loss - loss_o.item()) < 1e-5 * max(1.0, abs(loss_o.item()))
    errs = {(l, k): frob_rel(leaves[l][k].grad.cpu(), o_levels[l][k].grad) for l in range(4) for k in ("x0", "x1", "s0", "s1")}
    errs["R0"], errs["t0"] = frob_rel(gR, Ro.grad), frob_rel(gt, to.grad)
    print("full-resolution gradient errors:", {k: f"{v:.1e}" for k, v in errs.items()})
    assert max(errs.values()) < 5e-4, errs
    assert max(v for k, v in errs.items() if k[0] == 0 or isinstance(k, str)) < 1e-4, errs      # finest level and pose: <= 1e-5 measured


def test_gradients_are_reproducible_to_rounding():
    """The bilinear adjoint scatters with atomicAdd, so the order of the fp32 additions varies from run to run: the
    spread is rounding noise (asserted <= 1e-5 Frobenius-relative per tensor), not bit-identical -- documented here."""
    data, R0, t0, cs = _full_size_case(B=3, seed=92)

    def loss_cuda(outs):
        return sum((R * c[0].to(DEV)).sum() + (t * c[1].to(DEV)).sum() for (R, t, _), c in zip(outs, cs))

    runs = []
    for _ in range(3):
        _, leaves, gR, gt, _ = cuda_grads(data["levels"], R0, t0, 3, True, loss_cuda)
        runs.append(([leaves[l][k].grad.clone() for l in range(4) for k in ("x0", "x1", "s0", "s1")], gR, gt))
    worst, identical = 0.0, True
    for r in runs[1:]:
        for a, b in zip(r[0] + [r[1], r[2]], runs[0][0] + [runs[0][1], runs[0][2]]):
            worst = max(worst, frob_rel(a, b))
            identical = identical and torch.equal(a, b)
    print(f"run-to-run gradient spread: {worst:.1e} (bit-identical: {identical})")
    assert worst < 1e-5, worst
