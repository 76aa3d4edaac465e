"""Access to the installed reference (baseline/_ref, see install_reference.py): module handles, a
LeastSquareTracking built with the flags of the reference's own scripts, synthetic RGB-D pairs.

Only tests/, bench.py --impl reference and __graft_entry__ import this; the product never does."""
from __future__ import annotations

import argparse
import contextlib
import importlib
import io
import os
import sys
from typing import List, Optional

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
CODE = os.path.join(HERE, "_ref", "code")

# scripts/eval_tum_rgbd.sh / train_tum_rgbd.sh of the reference (TUM, U_IC): 8 feature channels, one uncertainty channel,
# laplacian uncertainty, remove_tru_sigma, constant ICP scaler
EVAL_TUM_FLAGS = ["--encoder_name", "ConvRGBD2", "--mestimator", "None", "--solver", "Direct-Nodamping",
                  "--feature_channel", "8", "--uncertainty_channel", "1", "--feature_extract", "conv",
                  "--uncertainty", "laplacian", "--remove_tru_sigma", "--scaler", "None"]


# scripts/train_tum_rgbd.sh: the same tracker with the learned initial pose (SFMPoseNet) trained jointly
TRAIN_TUM_FLAGS = EVAL_TUM_FLAGS + ["--init_pose", "sfm_net", "--train_init_pose", "--multi_hypo", "prob_fuse"]


def available() -> bool:
    return os.path.isdir(CODE)


def modules():
    """(algorithms, geometry, LeastSquareTracking module, config) of the installed reference."""
    if not available():
        raise RuntimeError("baseline/_ref is missing: run `python baseline/install_reference.py` in the build container")
    if CODE not in sys.path:
        sys.path.insert(0, CODE)
    import warnings
    warnings.filterwarnings("ignore", category=SyntaxWarning)     # 2019 docstrings with LaTeX backslashes
    alg = importlib.import_module("models.algorithms")
    geo = importlib.import_module("models.geometry")
    lst = importlib.import_module("models.LeastSquareTracking")
    cfg = importlib.import_module("config")
    return alg, geo, lst, cfg


def options(flags: Optional[List[str]] = None):
    _, _, _, cfg = modules()
    parser = argparse.ArgumentParser()
    cfg.add_tracking_config(parser)
    cfg.add_basics_config(parser)
    cfg.add_test_basics_config(parser)
    return parser.parse_args(EVAL_TUM_FLAGS if flags is None else flags)


def make_tracker(flags: Optional[List[str]] = None, seed: int = 0):
    """LeastSquareTracking as evaluate.py / train.py build it (random weights: the checkpoints are not shipped)."""
    _, _, lst, _ = modules()
    opt = options(flags)
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):   # the constructor prints its configuration
        net = lst.LeastSquareTracking(encoder_name=opt.encoder_name, uncertainty_type=opt.uncertainty,
                                      direction=opt.direction, max_iter_per_pyr=opt.max_iter_per_pyr,
                                      mEst_type=opt.mestimator, solver_type=opt.solver, tr_samples=opt.tr_samples,
                                      options=opt, no_weight_sharing=opt.no_weight_sharing)
    return net


def synthetic_rgbd(B: int, H: int, W: int, seed: int = 0, device="cpu"):
    """Smooth random RGB-D frame pairs with a small relative motion: (img0, img1, depth0, depth1, K)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(seed)

    def smooth(c):
        low = torch.randn((B, c, max(H // 8, 2), max(W // 8, 2)), generator=g)
        return F.interpolate(low, (H, W), mode="bilinear", align_corners=True)

    img0 = (0.5 + 0.25 * smooth(3)).clamp(0, 1)
    depth0 = (1.5 + 0.3 * smooth(1)).clamp(0.5, 5.0)
    shift = 2
    img1 = torch.roll(img0, shift, 3) + 0.01 * torch.randn((B, 3, H, W), generator=g)
    depth1 = torch.roll(depth0, shift, 3)
    K = torch.tensor([525.0, 525.0, 319.5, 239.5]) * (W / 640.0)
    K = K.repeat(B, 1)
    return tuple(x.to(device) for x in (img0, img1.clamp(0, 1), depth0, depth1, K))


def solver_chain(net, levels, pose):
    """The coarse-to-fine chain LeastSquareTracking.forward runs through tr_update3..0 (LeastSquareTracking.py:345-446)
    on precomputed per-level tensors: ``levels`` coarse first, each with x0, x1, s0, s1, invD0, invD1, K (K already
    scaled to the level).  Returns the final [R, t]."""
    updates = [net.tr_update3, net.tr_update2, net.tr_update1, net.tr_update0][-len(levels):]
    prior = None
    for tr, lv in zip(updates, levels):
        out = tr(pose, lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"], lv["s0"], lv["s1"], wPrior=prior,
                 depth0=lv.get("depth0"), depth1=lv.get("depth1"), vis_res=False)
        pose, prior = out[0], out[1]
    return pose
