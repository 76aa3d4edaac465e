"""The drop-in against the REAL reference: a LeastSquareTracking built by the reference's own code (baseline/_ref,
installed by baseline/install_reference.py) runs on CUDA next to a copy whose tr_update0..3 were swapped by
patch_tracker -- same weights, same RGB-D input through the reference's own feature encoder.

  eval   poses of the patched tracker == the reference's (the reference's CUDA grid_sample / bmm against this
         implementation's kernels): <= 1e-5 on R and t
  train  one training step as train.py runs it (train.py:117-192: forward in train mode, loss on the pose pyramid,
         backward): the pose pyramid and the gradients of every encoder parameter agree
"""
import copy

import pytest
import torch

from baseline import reference as REF
from deep_prob_feature_track_b200 import algorithms as A
from helpers import frob_rel

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not REF.available(), reason="baseline/_ref not installed")]
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def exact_convolutions():
    """Both trackers share the encoder; keep cuDNN / cuBLAS off TF32 so the reference's own bmm is fp32 too."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def trackers(flags=None):
    ref = REF.make_tracker(flags, seed=0).to(DEV)
    ours = A.patch_tracker(copy.deepcopy(ref))
    return ref, ours


def test_eval_forward_matches_the_reference_on_cuda():
    ref, ours = trackers()
    ref.eval(), ours.eval()
    for i in range(4):
        assert type(getattr(ours, f"tr_update{i}")).__module__.startswith("deep_prob_feature_track_b200")
        assert type(getattr(ref, f"tr_update{i}")).__module__ == "models.algorithms"
    img0, img1, d0, d1, K = REF.synthetic_rgbd(4, 120, 160, seed=3, device=DEV)
    with torch.no_grad():
        R_ref, t_ref = ref(img0, img1, d0, d1, K)
        R, t = ours(img0, img1, d0, d1, K)
    assert R.shape == R_ref.shape and t.shape == t_ref.shape
    assert t_ref.abs().max() > 1e-5                       # the solver moved the pose (random weights: a small motion)
    assert (R - R_ref).abs().max() < 1e-5, (R - R_ref).abs().max()
    assert (t - t_ref).abs().max() < 1e-5, (t - t_ref).abs().max()


def test_icp_variant_matches_the_reference_on_cuda():
    """train_tum_feature_icp.sh: the same tracker with --combine_ICP (constant scaler)."""
    ref, ours = trackers(REF.EVAL_TUM_FLAGS + ["--combine_ICP"])
    ref.eval(), ours.eval()
    img0, img1, d0, d1, K = REF.synthetic_rgbd(2, 120, 160, seed=5, device=DEV)
    with torch.no_grad():
        R_ref, t_ref = ref(img0, img1, d0, d1, K)
        R, t = ours(img0, img1, d0, d1, K)
    assert (R - R_ref).abs().max() < 1e-5 and (t - t_ref).abs().max() < 1e-5


def test_training_step_gradients_match_the_reference():
    alg, geo, lst, cfg = REF.modules()
    import models.criterions as crit        # the reference's loss (criterions.py:101-136)
    ref, ours = trackers()
    ref.train(), ours.train()
    B = 4
    img0, img1, d0, d1, K = REF.synthetic_rgbd(B, 120, 160, seed=7, device=DEV)
    R_gt = torch.eye(3, device=DEV).repeat(B, 1, 1)
    t_gt = torch.tensor([[0.01, -0.005, 0.002]], device=DEV).repeat(B, 1)
    invalid = (d0 < 0.1)

    def step(net):
        net.zero_grad()
        Rs, ts = net(img0, img1, d0, d1, K)                      # (B, N, 3, 3), (B, N, 3)
        loss = crit.compute_RT_EPE_loss(Rs, ts, R_gt, t_gt, d0, K, invalid=invalid).mean() * 1e2     # train.py:168
        loss.backward()
        return Rs.detach(), ts.detach(), loss.detach(), {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}

    Rr, tr_, lr, gr = step(ref)
    Ro, to, lo, go = step(ours)
    assert (Ro - Rr).abs().max() < 1e-5 and (to - tr_).abs().max() < 1e-5
    assert abs(lo.item() - lr.item()) <= 1e-5 * max(1.0, abs(lr.item()))
    assert set(go) == set(gr) and len(gr) > 20
    worst = max(frob_rel(go[n], gr[n]) for n in gr if gr[n].norm() > 1e-12)
    assert worst < 2e-3, worst
    total = frob_rel(torch.cat([go[n].flatten() for n in sorted(gr)]), torch.cat([gr[n].flatten() for n in sorted(gr)]))
    assert total < 5e-4, total


# run_example.py:84-92 (BASELINE config 1): the DeepIC tracker -- IC solver, convolutional M-estimator, residual-volume damping
DEEPIC_FLAGS = ["--encoder_name", "ConvRGBD2", "--mestimator", "MultiScale2w", "--solver", "Direct-ResVol", "--uncertainty", "None"]


def test_deepic_run_example_configuration_matches_the_reference_on_cuda():
    ref, ours = trackers(DEEPIC_FLAGS)
    ref.eval(), ours.eval()
    for i in range(4):
        assert type(getattr(ours, f"tr_update{i}")).__name__ == "TrustRegionBase"
        assert type(getattr(ours, f"tr_update{i}")).__module__.startswith("deep_prob_feature_track_b200")
    for B, seed in ((1, 11), (3, 12)):                    # run_example.py feeds one pair at a time
        img0, img1, d0, d1, K = REF.synthetic_rgbd(B, 120, 160, seed=seed, device=DEV)
        with torch.no_grad():
            R_ref, t_ref = ref(img0, img1, d0, d1, K)
            R, t = ours(img0, img1, d0, d1, K)
        assert t_ref.abs().max() > 1e-6
        assert (R - R_ref).abs().max() < 2e-5, (R - R_ref).abs().max()
        assert (t - t_ref).abs().max() < 2e-5, (t - t_ref).abs().max()


def test_deepic_training_step_gradients_match_the_reference():
    """One training step of the DeepIC tracker: autograd chains the dpft_ic_* functions around the reference's own
    M-estimator CNN and damping MLP; gradients of every parameter (encoder, M-estimator, solver MLP) agree."""
    import models.criterions as crit
    ref, ours = trackers(DEEPIC_FLAGS)
    ref.train(), ours.train()
    B = 2
    img0, img1, d0, d1, K = REF.synthetic_rgbd(B, 120, 160, seed=13, device=DEV)
    R_gt = torch.eye(3, device=DEV).repeat(B, 1, 1)
    t_gt = torch.tensor([[0.01, -0.005, 0.002]], device=DEV).repeat(B, 1)
    invalid = (d0 < 0.1)

    def step(net):
        net.zero_grad()
        Rs, ts = net(img0, img1, d0, d1, K)
        loss = crit.compute_RT_EPE_loss(Rs, ts, R_gt, t_gt, d0, K, invalid=invalid).mean() * 1e2
        loss.backward()
        return Rs.detach(), ts.detach(), {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}

    Rr, tr_, gr = step(ref)
    Ro, to, go = step(ours)
    assert (Ro - Rr).abs().max() < 2e-5 and (to - tr_).abs().max() < 2e-5
    assert set(go) == set(gr) and len(gr) > 20
    total = frob_rel(torch.cat([go[n].flatten() for n in sorted(gr)]), torch.cat([gr[n].flatten() for n in sorted(gr)]))
    assert total < 2e-3, total


def test_fused_eval_forward_equals_the_level_by_level_forward():
    """patch_tracker(net, fused_forward=True): the reference's _preprocess, then all four levels in ONE solver call.
    Same kernels on the same tensors as the four module calls, so the poses are the same floats; whatever the one call
    does not serve (here: train mode) takes the original forward."""
    ref, ours = trackers()
    fused = A.patch_tracker(copy.deepcopy(ref), fused_forward=True)
    ref.eval(), ours.eval(), fused.eval()
    img0, img1, d0, d1, K = REF.synthetic_rgbd(4, 120, 160, seed=21, device=DEV)
    with torch.no_grad():
        R_ref, t_ref = ref(img0, img1, d0, d1, K)
        R_a, t_a = ours(img0, img1, d0, d1, K)
        R_b, t_b = fused(img0, img1, d0, d1, K)
    assert R_b.shape == R_a.shape and t_b.shape == t_a.shape
    assert torch.equal(R_a, R_b) and torch.equal(t_a, t_b)
    assert (R_b - R_ref).abs().max() < 1e-5 and (t_b - t_ref).abs().max() < 1e-5
    fused.train(), ours.train()
    Rs_a, ts_a = ours(img0, img1, d0, d1, K)
    Rs_b, ts_b = fused(img0, img1, d0, d1, K)                # falls back: the pose pyramid of train.py
    assert Rs_b.shape == Rs_a.shape and torch.allclose(Rs_a, Rs_b, atol=1e-6) and torch.allclose(ts_a, ts_b, atol=1e-6)
