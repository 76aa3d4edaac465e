"""compute_RT_EPE_loss drop-in (criterions.py:101-136) against the reference fixture and the oracle."""
import pytest
import torch

from deep_prob_feature_track_b200 import criterions as C
from helpers import frob_rel, load_golden
from oracle import ic_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_training_and_evaluation_calls_match_the_reference():
    g = load_golden("pose_loss")
    d = {k: v.to(DEV) for k, v in g.items()}
    R_est = d["R_est"].clone().requires_grad_(True)
    t_est = d["t_est"].clone().requires_grad_(True)
    loss = C.compute_RT_EPE_loss(R_est, t_est, d["R_gt"], d["t_gt"], d["depth"], d["K"], invalid=d["invalid"].bool())
    (loss * d["w"]).sum().backward()
    assert frob_rel(loss.cpu(), g["loss"]) < 1e-5, (loss.cpu(), g["loss"])
    assert float(loss[-1].detach()) == 0.0                      # the sample without a valid pixel
    assert frob_rel(R_est.grad.cpu(), g["g_R_est"]) < 1e-4 and frob_rel(t_est.grad.cpu(), g["g_t_est"]) < 1e-4
    with torch.no_grad():
        ev = C.compute_RT_EPE_loss(d["R_est"][:, 0], d["t_est"][:, 0], d["R_gt"], d["t_gt"], d["depth"], d["K"],
                                   invalid=d["invalid"].bool())
    assert frob_rel(ev.cpu(), g["loss_eval"]) < 1e-5


@pytest.mark.parametrize("B,N,h,w", [(1, 1, 7, 5), (5, 5, 60, 80), (64, 4, 60, 80), (2, 8, 33, 47)])
def test_against_the_oracle_without_a_mask(B, N, h, w):
    gen = torch.Generator().manual_seed(B * 100 + N)
    depth = torch.rand((B, 1, h, w), generator=gen) + 0.5
    K = torch.tensor([[60.0, 61.0, w / 2.0, h / 2.0]]).repeat(B, 1)
    from deep_prob_feature_track_b200.synthetic import _twist_to_pose
    R_gt, t_gt = _twist_to_pose((torch.rand((B, 6), generator=gen) - 0.5) * 0.2)
    est = [_twist_to_pose((torch.rand((B, 6), generator=gen) - 0.5) * 0.2) for _ in range(N)]
    R_est = torch.stack([e[0] for e in est], 1)
    t_est = torch.stack([e[1] for e in est], 1)
    R_est[:, 0], t_est[:, 0] = R_gt, t_gt                                    # zero distance: zero gradient, no NaN
    wgt = torch.rand((B,), generator=gen) + 0.5
    Ro, to = R_est.clone().requires_grad_(True), t_est.clone().requires_grad_(True)
    lo = O.pose_epe_loss(Ro, to, R_gt, t_gt, depth, K, None)
    (lo * wgt).sum().backward()
    Rc, tc = R_est.to(DEV).requires_grad_(True), t_est.to(DEV).requires_grad_(True)
    lc = C._PoseEpeFn.apply(depth.to(DEV), None, K.to(DEV), R_gt.to(DEV), t_gt.to(DEV), Rc, tc)
    (lc * wgt.to(DEV)).sum().backward()
    assert (lc.cpu() - lo.detach()).abs().max() < 1e-5 * (1.0 + lo.detach().abs().max())
    assert torch.isfinite(Rc.grad).all() and torch.isfinite(tc.grad).all()
    assert (Rc.grad.cpu() - Ro.grad).abs().max() < 1e-4 * (1.0 + Ro.grad.abs().max())
    assert (tc.grad.cpu() - to.grad).abs().max() < 1e-4 * (1.0 + to.grad.abs().max())


def test_nan_target_points_are_skipped():
    """NaN depth -> NaN target point -> not counted (criterions.py:31-33).  Forward only: the reference's autograd
    turns such pixels into NaN gradients, this backward leaves them out."""
    gen = torch.Generator().manual_seed(3)
    B, N, h, w = 3, 2, 20, 30
    depth = torch.rand((B, 1, h, w), generator=gen) + 0.5
    depth[torch.rand((B, 1, h, w), generator=gen) < 0.2] = float("nan")
    K = torch.tensor([[40.0, 41.0, w / 2.0, h / 2.0]]).repeat(B, 1)
    from deep_prob_feature_track_b200.synthetic import _twist_to_pose
    R_gt, t_gt = _twist_to_pose((torch.rand((B, 6), generator=gen) - 0.5) * 0.2)
    est = [_twist_to_pose((torch.rand((B, 6), generator=gen) - 0.5) * 0.2) for _ in range(N)]
    R_est, t_est = torch.stack([e[0] for e in est], 1), torch.stack([e[1] for e in est], 1)
    lo = O.pose_epe_loss(R_est, t_est, R_gt, t_gt, depth, K, None)
    Rc, tc = R_est.to(DEV).requires_grad_(True), t_est.to(DEV).requires_grad_(True)
    lc = C._PoseEpeFn.apply(depth.to(DEV), None, K.to(DEV), R_gt.to(DEV), t_gt.to(DEV), Rc, tc)
    lc.sum().backward()
    assert torch.isfinite(lo).all() and frob_rel(lc.detach().cpu(), lo) < 1e-5
    assert torch.isfinite(Rc.grad).all() and torch.isfinite(tc.grad).all()


def test_cpu_tensors_are_rejected():
    g = load_golden("pose_loss")
    with pytest.raises(RuntimeError):
        C.compute_RT_EPE_loss(g["R_est"], g["t_est"], g["R_gt"], g["t_gt"], g["depth"], g["K"], invalid=g["invalid"].bool())
