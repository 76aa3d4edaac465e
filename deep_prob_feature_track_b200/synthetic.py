"""Seeded synthetic RGB-D-like frame pairs at solver level.

There are no datasets or checkpoints in this environment, so the benchmark, the
smoke test and the parity tests all run on what this module makes: band-limited
feature maps, a smooth uncertainty field and a smooth depth map for the live frame,
and a keyframe obtained by resampling the live frame under a small ground-truth
motion, so that the Gauss-Newton iterations have something to converge to.

Everything is generated on the CPU with a ``torch.Generator`` so that the same seed
gives the same bytes on every box.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

TUM_K = (525.0, 525.0, 319.5, 239.5)   # intrinsics of a 640-wide TUM frame


def _smooth(gen: torch.Generator, shape: Tuple[int, int, int, int], coarse: int = 8) -> torch.Tensor:
    """N(0,1) noise at 1/coarse resolution, bilinearly upsampled: band-limited, unit-ish scale."""
    B, C, H, W = shape
    h, w = max(2, H // coarse + 1), max(2, W // coarse + 1)
    z = torch.randn((B, C, h, w), generator=gen)
    return F.interpolate(z, size=(H, W), mode="bilinear", align_corners=True)


def _twist_to_pose(xi: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Rodrigues for the rotation part, translation taken as is. xi: (B,6) [rot, trs]."""
    w = xi[:, :3]
    theta = w.norm(dim=1).clamp_min(1e-12).view(-1, 1, 1)
    k = w / theta.view(-1, 1)
    Kx = torch.zeros(xi.shape[0], 3, 3)
    Kx[:, 0, 1], Kx[:, 0, 2] = -k[:, 2], k[:, 1]
    Kx[:, 1, 0], Kx[:, 1, 2] = k[:, 2], -k[:, 0]
    Kx[:, 2, 0], Kx[:, 2, 1] = -k[:, 1], k[:, 0]
    R = torch.eye(3).expand_as(Kx) + Kx * torch.sin(theta) + Kx.bmm(Kx) * (1 - torch.cos(theta))
    return R, xi[:, 3:].clone()


def _warp_grid(depth0: torch.Tensor, R: torch.Tensor, t: torch.Tensor, K: torch.Tensor):
    """Pixel coordinates and depth of every keyframe pixel seen from the live frame."""
    B, _, H, W = depth0.shape
    fx, fy, cx, cy = (K[:, i].view(B, 1, 1) for i in range(4))
    cols = torch.arange(W, dtype=torch.float32).view(1, 1, W)
    rows = torch.arange(H, dtype=torch.float32).view(1, H, 1)
    x = ((cols - cx) / fx).expand(B, H, W)
    y = ((rows - cy) / fy).expand(B, H, W)
    z = depth0[:, 0]
    P = torch.stack((x * z, y * z, z), dim=1).view(B, 3, -1)
    Q = (R.bmm(P) + t.view(B, 3, 1)).view(B, 3, H, W)
    u = Q[:, 0] / Q[:, 2] * fx + cx
    v = Q[:, 1] / Q[:, 2] * fy + cy
    return u, v, Q[:, 2]


def _lookup(img: torch.Tensor, u: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    B, C, H, W = img.shape
    grid = torch.stack((u / ((W - 1) / 2) - 1, v / ((H - 1) / 2) - 1), dim=3)
    return F.grid_sample(img, grid, mode="bilinear", padding_mode="border", align_corners=True)


def make_frame_pairs(B: int, C: int, H: int, W: int, *, seed: int = 1234, n_levels: int = 4,
                     motion: float = 0.02, sigma_channels: int = 1, with_depth: bool = False,
                     noise: float = 0.01) -> Dict:
    """Build ``B`` frame pairs and their ``n_levels``-level pyramids.

    Returns a dict with
      ``levels``  list ordered COARSE to FINE (the order the solver consumes them); each entry
                  has x0, x1, s0, s1 (B,C,h,w), invD0, invD1 (B,1,h,w), K (B,4)
                  [, depth0, depth1 when ``with_depth``];
      ``R_gt``, ``t_gt``  the motion the keyframe was resampled with;
      ``R0``, ``t0``      the identity initial pose (B,3,3), (B,3).
    Feature/uncertainty pyramids are average-pooled, depth pyramids max-pooled, like the
    reference's ImagePyramids (algorithms.py:1201-1219); inverse depth is clamp(1/d,0,10)
    with the batch-global extremes zeroed (LeastSquareTracking.py:656-661).
    """
    gen = torch.Generator().manual_seed(seed)
    K0 = torch.tensor(TUM_K, dtype=torch.float32) * (W / 640.0)
    K = K0.view(1, 4).repeat(B, 1)

    depth1 = (1.5 + 0.3 * _smooth(gen, (B, 1, H, W))).clamp(0.5, 5.0)
    x1 = _smooth(gen, (B, C, H, W), coarse=4) + 0.25 * _smooth(gen, (B, C, H, W), coarse=2)
    s1 = torch.exp((0.3 * _smooth(gen, (B, sigma_channels, H, W))).clamp(-3, 3))

    xi = (torch.rand((B, 6), generator=gen) * 2 - 1) * motion
    R_gt, t_gt = _twist_to_pose(xi)

    # keyframe depth consistent with the live depth under the motion (fixed point, 3 rounds)
    depth0 = depth1.clone()
    for _ in range(3):
        u, v, z_pred = _warp_grid(depth0, R_gt, t_gt, K)
        depth0 = (depth0 - (z_pred.unsqueeze(1) - _lookup(depth1, u, v))).clamp(0.5, 5.0)
    u, v, _ = _warp_grid(depth0, R_gt, t_gt, K)
    x0 = _lookup(x1, u, v) + noise * torch.randn((B, C, H, W), generator=gen)
    s0 = _lookup(s1, u, v) * torch.exp(noise * torch.randn((B, sigma_channels, H, W), generator=gen))

    # ~2 % of the depth pixels sit on the clip bounds (invalid in a real sensor)
    for d in (depth0, depth1):
        hole = torch.rand((B, 1, H, W), generator=gen)
        d[hole < 0.01] = 0.5
        d[hole > 0.99] = 5.0

    def inv_depth(d):
        inv = torch.clamp(1.0 / d, 0, 10)
        lo, hi = inv.min(), inv.max()
        inv = torch.where((inv == lo) | (inv == hi), torch.zeros_like(inv), inv)
        return inv

    invD0, invD1 = inv_depth(depth0), inv_depth(depth1)
    if sigma_channels != C:
        s0 = s0.repeat(1, C // sigma_channels, 1, 1)
        s1 = s1.repeat(1, C // sigma_channels, 1, 1)

    levels: List[Dict] = []
    for l in range(n_levels - 1, -1, -1):
        k = 1 << l
        lv = {
            "x0": F.avg_pool2d(x0, k, k).contiguous(), "x1": F.avg_pool2d(x1, k, k).contiguous(),
            "s0": F.avg_pool2d(s0, k, k).contiguous(), "s1": F.avg_pool2d(s1, k, k).contiguous(),
            "invD0": F.max_pool2d(invD0, k, k).contiguous(), "invD1": F.max_pool2d(invD1, k, k).contiguous(),
            "K": (K / float(k)).contiguous(),
        }
        if with_depth:
            lv["depth0"] = F.max_pool2d(depth0, k, k).contiguous()
            lv["depth1"] = F.max_pool2d(depth1, k, k).contiguous()
        levels.append(lv)
    return {
        "levels": levels, "R_gt": R_gt, "t_gt": t_gt,
        "R0": torch.eye(3).repeat(B, 1, 1), "t0": torch.zeros(B, 3),
    }


def levels_to(levels: List[Dict], device, non_blocking: bool = False) -> List[Dict]:
    return [{k: v.to(device, non_blocking=non_blocking) for k, v in lv.items()} for lv in levels]
