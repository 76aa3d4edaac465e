"""Depth stage of _preprocess against the reference's own op chain restated with torch (bit-exact)."""
import pytest
import torch
import torch.nn.functional as F

from deep_prob_feature_track_b200 import algorithms as A

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def reference_chain(depth, n_levels):
    """LeastSquareTracking.py:656-661 + ImagePyramids (alg:1201-1219), verbatim semantics."""
    inv = torch.clamp(1.0 / depth, 0, 10)
    inv[inv == inv.min()] = 0
    inv[inv == inv.max()] = 0
    return [F.max_pool2d(inv, 1 << l, 1 << l) for l in range(n_levels)], [F.max_pool2d(depth, 1 << l, 1 << l) for l in range(n_levels)]


@pytest.mark.parametrize("B,H,W,n_levels", [(3, 120, 160, 4), (2, 37, 53, 3), (1, 8, 8, 4), (2, 480, 640, 4)])
def test_depth_pyramids_bit_exact(B, H, W, n_levels):
    g = torch.Generator().manual_seed(H)
    depth = (torch.rand((B, 1, H, W), generator=g) * 5 + 0.3).clamp(0.5, 5.0)     # clip bounds are hit -> zeroed
    depth[0, 0, 0, 0] = 0.05                                                        # 1/d above the clamp
    inv_ref, dpt_ref = reference_chain(depth.clone(), n_levels)
    inv, dpt = A.depth_pyramids(depth.to(DEV), n_levels, with_depth=True)
    for l in range(n_levels):
        assert torch.equal(inv[l].cpu(), inv_ref[l]), l
        assert torch.equal(dpt[l].cpu(), dpt_ref[l]), l
    assert (inv_ref[0] == 0).float().mean() > 0.01


def test_constant_depth_is_all_invalid():
    depth = torch.full((2, 1, 16, 16), 2.0)
    inv_ref, _ = reference_chain(depth.clone(), 2)
    inv = A.depth_pyramids(depth.to(DEV), 2)
    assert torch.equal(inv[0].cpu(), inv_ref[0]) and float(inv[0].abs().max()) == 0.0
