"""Generate the golden fixtures in this directory by running the REFERENCE itself.

Run in the build container only (it needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py

The reference is imported from where it lies -- nothing is copied into the repo.
Two run-time shims make the 2019-era code run on torch 2.x:
  * ``K >> n`` on a float tensor (LeastSquareTracking.py:350,374,398) -> K / 2**n;
    the chain below divides explicitly instead of going through LeastSquareTracking;
  * ``Tensor.split`` hands out clones while a reference module runs, so the in-place
    ``squeeze_`` on its outputs (algorithms.py:873-875) is legal under autograd.

Every fixture stores its inputs, so the tests never depend on RNG reproducibility.
Masks are stored as uint8.
"""
from __future__ import annotations

import contextlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/code")

import models.algorithms as alg  # noqa: E402  (the reference)

from deep_prob_feature_track_b200.synthetic import make_frame_pairs  # noqa: E402


@contextlib.contextmanager
def split_clones():
    orig = torch.Tensor.split

    def split(self, *a, **k):
        return tuple(p.clone() for p in orig(self, *a, **k))

    torch.Tensor.split = split
    try:
        yield
    finally:
        torch.Tensor.split = orig


class Recorder:
    """Capture occ / JtWJ / JtR / start pose of every iteration of a reference U_IC module."""

    def __init__(self, module):
        self.rows = []
        self._occ = None
        self._orig_res = alg.compute_inverse_residuals
        self._orig_gn = module.GN_solver
        self.module = module

    def __enter__(self):
        def res_hook(*a, **k):
            out = self._orig_res(*a, **k)
            self._occ = out[3].detach().clone()
            return out

        def gn_hook(JtJ, JtR, pose0, **k):
            self.rows.append(dict(A=JtJ.detach().clone(), b=JtR.detach().clone(),
                                  R=pose0[0].detach().clone(), t=pose0[1].detach().reshape(-1, 3).clone(),
                                  occ=self._occ))
            return self._orig_gn(JtJ, JtR, pose0, **k)

        alg.compute_inverse_residuals = res_hook
        self.module.GN_solver = gn_hook
        return self

    def __exit__(self, *exc):
        alg.compute_inverse_residuals = self._orig_res
        del self.module.GN_solver

    def stacked(self):
        return {k: torch.stack([r[k] for r in self.rows]).numpy() for k in ("A", "b", "R", "t")} | {
            "occ": torch.stack([r["occ"] for r in self.rows]).to(torch.uint8).numpy()}


def perturbed_pose(B, seed, scale=0.01):
    g = torch.Generator().manual_seed(seed)
    xi = (torch.rand((B, 6), generator=g) * 2 - 1) * scale
    from deep_prob_feature_track_b200.synthetic import _twist_to_pose
    return _twist_to_pose(xi)


def np_level(lv):
    return {k: v.numpy() for k, v in lv.items()}


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def uic_single_level(name, *, seed, B=2, C=4, H=24, W=32, remove_tru_sigma=False, combine_icp=False,
                     masks=False, iters=3):
    data = make_frame_pairs(B, C, H, W, seed=seed, n_levels=1, with_depth=combine_icp)
    lv = data["levels"][0]
    R0, t0 = perturbed_pose(B, seed + 1)
    scale = alg.ScaleNet("None") if combine_icp else None
    mod = alg.TrustRegionInverseWUncertainty(max_iter=iters, combine_icp=combine_icp, scale_func=scale,
                                             remove_tru_sigma=remove_tru_sigma, uncer_prop=True)
    extra = {}
    kw = {}
    if masks:
        g = torch.Generator().manual_seed(seed + 2)
        kw["obj_mask0"] = torch.rand((B, 1, H, W), generator=g) > 0.2
        kw["obj_mask1"] = torch.rand((B, 1, H, W), generator=g) > 0.2
        extra = {k: v.to(torch.uint8).numpy() for k, v in kw.items()}
    with torch.no_grad(), Recorder(mod) as rec:
        pose, weights, A_last = mod(
            [R0, t0], lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"], lv["s0"], lv["s1"],
            wPrior=None, depth0=lv.get("depth0"), depth1=lv.get("depth1"), vis_res=False, **kw)
        loss = mod.forward_residuals(
            [R0, t0], lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"], lv["s0"], lv["s1"],
            wPrior=None, depth0=lv.get("depth0"), depth1=lv.get("depth1"), vis_res=False, **kw)
    hist = rec.stacked()
    save(name, **{f"in_{k}": v for k, v in np_level(lv).items()}, **extra,
         R0=R0.numpy(), t0=t0.numpy(), R_out=pose[0].numpy(), t_out=pose[1].numpy(),
         weights=weights.numpy(), A_last=A_last.numpy(), res_loss=loss.numpy(),
         **{f"it_{k}": v for k, v in hist.items()},
         flags=np.array([remove_tru_sigma, combine_icp, masks, iters], dtype=np.int32))


def uic_pyramid(name, *, seed, B=2, C=4, H=48, W=64, remove_tru_sigma=True, iters=3):
    """Chain four reference modules exactly as LeastSquareTracking.forward does (:345-446)."""
    data = make_frame_pairs(B, C, H, W, seed=seed, n_levels=4)
    R, t = data["R0"], data["t0"].view(B, 3, 1)
    arrays = {}
    # identity start would divide by theta=0 only if xi==0, which noise in the data prevents
    with torch.no_grad():
        for i, lv in enumerate(data["levels"]):
            mod = alg.TrustRegionInverseWUncertainty(max_iter=iters, remove_tru_sigma=remove_tru_sigma)
            with Recorder(mod) as rec:
                (R, t), _ = mod([R, t], lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"], lv["s0"],
                                lv["s1"], vis_res=False)
            for k, v in np_level(lv).items():
                arrays[f"in{i}_{k}"] = v
            for k, v in rec.stacked().items():
                arrays[f"it{i}_{k}"] = v
            arrays[f"R_lvl{i}"] = R.numpy()
            arrays[f"t_lvl{i}"] = t.numpy()
    save(name, **arrays, R_gt=data["R_gt"].numpy(), t_gt=data["t_gt"].numpy(),
         flags=np.array([remove_tru_sigma, 0, 0, iters], dtype=np.int32))


def uic_gradients(name, *, seed, B=2, C=3, H=16, W=20, iters=2, remove_tru_sigma=False):
    """Reference autograd through one level: d(sum of a fixed linear functional of the pose)."""
    data = make_frame_pairs(B, C, H, W, seed=seed, n_levels=1)
    lv = data["levels"][0]
    R0, t0 = perturbed_pose(B, seed + 1)
    leaves = {k: lv[k].clone().requires_grad_(True) for k in ("x0", "x1", "s0", "s1")}
    R0 = R0.clone().requires_grad_(True)
    t0 = t0.clone().requires_grad_(True)
    g = torch.Generator().manual_seed(seed + 3)
    cR = torch.randn((B, 3, 3), generator=g)
    ct = torch.randn((B, 3), generator=g)
    cA = torch.randn((B, 6, 6), generator=g) * 1e-4
    mod = alg.TrustRegionInverseWUncertainty(max_iter=iters, remove_tru_sigma=remove_tru_sigma, uncer_prop=True)
    mod.train()
    with split_clones():
        (R, t), _, A = mod([R0, t0], leaves["x0"], leaves["x1"], lv["invD0"], lv["invD1"], lv["K"],
                           leaves["s0"], leaves["s1"], vis_res=False)
        loss = (R * cR).sum() + (t * ct).sum() + (A * cA).sum()
        loss.backward()
    save(name, **{f"in_{k}": v for k, v in np_level(lv).items()}, R0=R0.detach().numpy(), t0=t0.detach().numpy(),
         cR=cR.numpy(), ct=ct.numpy(), cA=cA.numpy(), loss=loss.detach().numpy(),
         R_out=R.detach().numpy(), t_out=t.detach().numpy(),
         **{f"g_{k}": v.grad.numpy() for k, v in leaves.items()}, g_R0=R0.grad.numpy(), g_t0=t0.grad.numpy(),
         flags=np.array([remove_tru_sigma, 0, 0, iters], dtype=np.int32))


def ic_single_level(name, *, seed, B=2, C=4, H=24, W=32, solver="Direct-Nodamping", mest="None", iters=3):
    if mest != "None":
        C = 1
    data = make_frame_pairs(B, C, H, W, seed=seed, n_levels=1)
    lv = data["levels"][0]
    R0, t0 = perturbed_pose(B, seed + 1)
    torch.manual_seed(seed + 5)
    mest_net = alg.DeepRobustEstimator(mest).eval()
    solver_net = alg.DirectSolverNet(solver, samples=10).eval()
    if solver == "Direct-ResVol":
        # xavier init leaves the last ReLU mostly dead; give it a live bias so damping is non-trivial
        with torch.no_grad():
            solver_net.net[-1][0].bias.fill_(0.05)
    mod = alg.TrustRegionBase(max_iter=iters, mEst_func=mest_net, solver_func=solver_net).eval()
    wprior = torch.ones(B, 1, max(H // 2, 1), max(W // 2, 1)) * 0.001
    hist = []
    orig = alg.inverse_update_pose

    def hook(Hm, rhs, pose):
        hist.append((Hm.detach().clone(), rhs.detach().clone()))
        return orig(Hm, rhs, pose)

    alg.inverse_update_pose = hook
    try:
        with torch.no_grad():
            pose, weights = mod([R0, t0], lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"], wPrior=wprior)
            loss = mod.forward_residuals([R0, t0], lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"],
                                         wPrior=wprior)
    finally:
        alg.inverse_update_pose = orig
    # with ResVol the hook also fires for the 10 trial solves; the real update is every 11th call
    stride = 11 if solver == "Direct-ResVol" else 1
    real = hist[stride - 1::stride]
    nets = {}
    for prefix, net in (("mest", mest_net), ("solver", solver_net)):
        for k, v in net.state_dict().items():
            nets[f"{prefix}__{k}"] = v.numpy()
    save(name, **{f"in_{k}": v for k, v in np_level(lv).items()}, R0=R0.numpy(), t0=t0.numpy(),
         wprior=wprior.numpy(), R_out=pose[0].numpy(), t_out=pose[1].numpy(), weights=weights.numpy(),
         res_loss=loss.numpy(), it_H=torch.stack([h for h, _ in real]).numpy(),
         it_b=torch.stack([b for _, b in real]).numpy(), **nets,
         flags=np.array([0, 0, 0, iters], dtype=np.int32))


def ic_gradients(name, *, seed, B=2, C=1, H=20, W=28, iters=2):
    """Reference autograd through one DeepIC level (conv M-estimator + residual-volume damping MLP, eval-mode
    batch norm): gradients of a fixed linear functional of the pose w.r.t. the maps, the start pose and the
    first layer of each network."""
    data = make_frame_pairs(B, C, H, W, seed=seed, n_levels=1)
    lv = data["levels"][0]
    R0, t0 = perturbed_pose(B, seed + 1)
    torch.manual_seed(seed + 5)
    mest_net = alg.DeepRobustEstimator("MultiScale2w").eval()
    solver_net = alg.DirectSolverNet("Direct-ResVol", samples=10).eval()
    with torch.no_grad():
        solver_net.net[-1][0].bias.fill_(0.05)
    mod = alg.TrustRegionBase(max_iter=iters, mEst_func=mest_net, solver_func=solver_net).eval()
    wprior = torch.ones(B, 1, max(H // 2, 1), max(W // 2, 1)) * 0.001
    leaves = {k: lv[k].clone().requires_grad_(True) for k in ("x0", "x1")}
    R0 = R0.clone().requires_grad_(True)
    t0 = t0.clone().requires_grad_(True)
    g = torch.Generator().manual_seed(seed + 3)
    cR = torch.randn((B, 3, 3), generator=g)
    ct = torch.randn((B, 3), generator=g)
    with split_clones():
        (R, t), weights = mod([R0, t0], leaves["x0"], leaves["x1"], lv["invD0"], lv["invD1"], lv["K"], wPrior=wprior)
        loss = (R * cR).sum() + (t * ct).sum()
        loss.backward()
    nets = {}
    for prefix, net in (("mest", mest_net), ("solver", solver_net)):
        for k, v in net.state_dict().items():
            nets[f"{prefix}__{k}"] = v.numpy()
    save(name, **{f"in_{k}": v for k, v in np_level(lv).items()}, R0=R0.detach().numpy(), t0=t0.detach().numpy(),
         wprior=wprior.numpy(), cR=cR.numpy(), ct=ct.numpy(), loss=loss.detach().numpy(),
         R_out=R.detach().numpy(), t_out=t.detach().numpy(),
         **{f"g_{k}": v.grad.numpy() for k, v in leaves.items()}, g_R0=R0.grad.numpy(), g_t0=t0.grad.numpy(),
         g_mest_conv0=mest_net.net[0][0].weight.grad.numpy(), g_solver_fc0=solver_net.net[0][0].weight.grad.numpy(),
         g_solver_fc2_bias=solver_net.net[2][0].bias.grad.numpy(), **nets,
         flags=np.array([0, 0, 0, iters], dtype=np.int32))


def pose_loss(name, *, seed, B=3, N=4, H=120, W=160):
    """Reference compute_RT_EPE_loss (criterions.py:101-136), training and evaluation calls, with autograd."""
    import models.criterions as crit
    g = torch.Generator().manual_seed(seed)
    depth = torch.rand((B, 1, H, W), generator=g) * 2.0 + 0.5
    invalid = torch.rand((B, 1, H, W), generator=g) < 0.15
    invalid[B - 1] = True                                   # one sample without a valid pixel
    K = torch.tensor([[131.25, 131.25, 79.875, 59.875]]).repeat(B, 1) * (W / 160.0)
    R_gt, t_gt = perturbed_pose(B, seed + 1, 0.05)
    R_est = torch.stack([perturbed_pose(B, seed + 2 + n, 0.05)[0] for n in range(N)], dim=1).requires_grad_(True)
    t_est = torch.stack([perturbed_pose(B, seed + 2 + n, 0.05)[1] for n in range(N)], dim=1).requires_grad_(True)
    w = torch.rand((B,), generator=g) + 0.5
    loss = crit.compute_RT_EPE_loss(R_est, t_est, R_gt, t_gt, depth, K, invalid=invalid)
    (loss * w).sum().backward()
    with torch.no_grad():
        loss_eval = crit.compute_RT_EPE_loss(R_est[:, 0], t_est[:, 0], R_gt, t_gt, depth, K, invalid=invalid)
        rdepth = torch.nn.functional.interpolate(depth, size=(60, 80), mode='bilinear')
        rinvalid = torch.nn.functional.interpolate(invalid.float(), size=(60, 80), mode='bilinear')
    save(name, depth=depth.numpy(), invalid=invalid.numpy().astype(np.uint8), K=K.numpy(), R_gt=R_gt.numpy(),
         t_gt=t_gt.numpy(), R_est=R_est.detach().numpy(), t_est=t_est.detach().numpy(), w=w.numpy(),
         loss=loss.detach().numpy(), loss_eval=loss_eval.numpy(), g_R_est=R_est.grad.numpy(), g_t_est=t_est.grad.numpy(),
         rdepth=rdepth.numpy(), rinvalid=rinvalid.numpy())


def main():
    torch.set_num_threads(4)
    uic_single_level("uic_plain", seed=11)
    uic_single_level("uic_trusigma", seed=12, remove_tru_sigma=True)
    uic_single_level("uic_icp", seed=13, combine_icp=True)
    uic_single_level("uic_masks", seed=14, masks=True, remove_tru_sigma=True)
    uic_single_level("uic_c8_wide", seed=15, C=8, H=30, W=70, remove_tru_sigma=True)
    uic_pyramid("uic_pyramid", seed=21)
    uic_gradients("uic_grad", seed=31)
    uic_gradients("uic_grad_trusigma", seed=32, remove_tru_sigma=True)
    ic_single_level("ic_plain", seed=41)
    ic_single_level("ic_resvol", seed=42, solver="Direct-ResVol")
    ic_single_level("ic_deepic", seed=43, solver="Direct-ResVol", mest="MultiScale2w")
    ic_gradients("ic_grad", seed=51)
    pose_loss("pose_loss", seed=61)


if __name__ == "__main__":
    main()
