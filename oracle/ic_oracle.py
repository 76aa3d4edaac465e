"""CPU restatement of the reference's trust-region inverse-compositional solver.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  This file is the checker for
the CUDA path and the CPU arm that ``bench.py`` times; the product never calls it.

What it restates (paths relative to /root/reference/code/models; ``alg`` =
algorithms.py, ``geo`` = geometry.py):

* ``pixel_rays``           geo:63-85     generate_xy_grid
* ``sobel_unit``           alg:1844-1865 feature_gradient
* ``warp_rows``            alg:1884-1917 compute_jacobian_warping
* ``project``              geo:291-323   batch_warp_inverse_depth
* ``sample_border``        geo:353-365   warp_features (+ torch grid_sampler_2d)
* ``occlusion``            geo:334-350   check_occ
* ``uic_residuals``        alg:1960-2015 compute_inverse_residuals / compose_residuals
* ``uic_jacobian``         alg:867-887, 1867-1882
* ``normal_equations``     alg:812-834   compute_jtj / compute_jtr
* ``damp`` / ``gn_update`` alg:2017-2103 lev_mar_H, invH, least_square_solve,
                           inverse_update_pose;  geo:105-123,146-185
* ``icp_term``             alg:916-997, 2148-2171; geo:1129-1136
* ``uic_level``            alg:611-723   TrustRegionInverseWUncertainty.forward
* ``uic_residual_loss``    alg:725-786, 2119-2137
* ``ic_level``             alg:45-121    TrustRegionBase.forward
* ``direct_solver``        alg:1604-1691 DirectSolverNet.forward
* ``track_pyramid``        LeastSquareTracking.py:345-446 (coarse-to-fine chain)

Parity pinning: the reference ships no golden vectors for this path (its only
test file prints finite differences), so this oracle is pinned
against the reference itself: ``tests/golden/make_golden.py`` imports the
reference from /root/reference in the build container, runs it on seeded
synthetic inputs and commits the outputs; ``tests/test_oracle_golden.py``
checks this file against them.

Two bilinear samplers are provided.  ``"grid_sample"`` calls torch's
``grid_sample`` exactly as the reference does (used for the CPU timing arm and
for the golden comparison).  ``"explicit"`` spells the upstream CUDA
``grid_sampler_2d`` arithmetic out as individually rounded fp32 operations in a
fixed order; it is the *definition* the CUDA kernels reproduce bit for bit on
everything that feeds a validity mask.

Everything is differentiable torch, so ``torch.autograd`` on this oracle is also
the oracle for the backward kernels.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Pose = Tuple[torch.Tensor, torch.Tensor]

EPS_UIC = 1e-6      # alg:2013  fill value of masked weighted residuals (U_IC)
EPS_IC = 1e-3       # alg:1955  fill value of masked residuals (IC)
OCC_THRES = 1e-1    # geo:334   z-buffer slack in inverse depth
ICP_DIST = 0.1      # alg:940   max point distance for an ICP correspondence


# --------------------------------------------------------------------------- geometry
def pixel_rays(K: torch.Tensor, H: int, W: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Normalised camera rays per pixel: x=(col-cx)/fx, y=(row-cy)/fy (geo:63-85)."""
    B = K.shape[0]
    fx, fy, cx, cy = (K[:, i].view(B, 1, 1, 1) for i in range(4))
    cols = torch.arange(W, dtype=K.dtype, device=K.device).view(1, 1, 1, W)
    rows = torch.arange(H, dtype=K.dtype, device=K.device).view(1, 1, H, 1)
    px = ((cols - cx) / fx).expand(B, 1, H, W)
    py = ((rows - cy) / fy).expand(B, 1, H, W)
    return px, py


def sobel_unit(img: torch.Tensor, normalise: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """Replicate-padded Sobel, optionally scaled to a unit 2-vector (alg:1844-1865)."""
    B, C, H, W = img.shape
    kx = torch.tensor([[-1., 0., 1.], [-2., 0., 2.], [-1., 0., 1.]], dtype=img.dtype,
                      device=img.device).view(1, 1, 3, 3)
    ky = kx.transpose(2, 3).contiguous()
    flat = F.pad(img.reshape(B * C, 1, H, W), (1, 1, 1, 1), mode="replicate")
    dx = F.conv2d(flat, kx)
    dy = F.conv2d(flat, ky)
    if normalise:
        mag = torch.sqrt(dx * dx + dy * dy + 1e-8)
        dx = dx / mag
        dy = dy / mag
    return dx.view(B, C, H, W), dy.view(B, C, H, W)


def warp_rows(invD: torch.Tensor, K: torch.Tensor, px: torch.Tensor, py: torch.Tensor):
    """d(u,v)/d(xi) at the identity, twist ordered [rot, trs] (alg:1884-1917)."""
    B = invD.shape[0]
    x = px.reshape(B, -1, 1)
    y = py.reshape(B, -1, 1)
    d = invD.reshape(B, -1, 1)
    xy = x * y
    zero = torch.zeros_like(d)
    du = torch.cat((-xy, 1 + x ** 2, -y, d, zero, -d * x), dim=2)
    dv = torch.cat((-1 - y ** 2, xy, x, zero, d, -d * y), dim=2)
    return du * K[:, 0].view(B, 1, 1), dv * K[:, 1].view(B, 1, 1)


def project(px, py, invD0, R, t, K):
    """SE(3) warp in inverse depth (geo:291-323).

    The rotation is written as ((r0*x + r1*y) + r2) + t*d with every product and
    sum rounded on its own: that order reproduces the reference's ``bmm`` result
    on CPU bit for bit and is what the CUDA kernel does.
    """
    B, _, H, W = px.shape
    t = t.reshape(B, 3)
    comps = []
    for i in range(3):
        r0, r1, r2 = (R[:, i, j].view(B, 1, 1, 1) for j in range(3))
        comps.append(((r0 * px + r1 * py) + r2) + t[:, i].view(B, 1, 1, 1) * invD0)
    wx, wy, wz = comps
    fx, fy, cx, cy = (K[:, i].view(B, 1, 1, 1) for i in range(4))
    u = (wx / wz) * fx + cx
    v = (wy / wz) * fy + cy
    inv_z = invD0 / wz
    return u, v, inv_z


def _unnormalise(coord: torch.Tensor, size: int) -> torch.Tensor:
    """warp_features' normalisation followed by grid_sampler's align_corners=True
    un-normalisation and border clip, kept as the same chain of fp32 ops."""
    g = coord / ((size - 1) / 2) - 1
    pix = ((g + 1) / 2) * (size - 1)
    return pix.clamp(0, size - 1)


def sample_border(img: torch.Tensor, u: torch.Tensor, v: torch.Tensor,
                  sampler: str = "explicit") -> torch.Tensor:
    """Bilinear lookup of ``img`` at pixel coordinates (u,v), border padding
    (geo:353-365).  u, v: (B,1,H,W) or (B,HW)."""
    B, C, H, W = img.shape
    u = u.reshape(B, H, W)
    v = v.reshape(B, H, W)
    if sampler == "grid_sample":
        grid = torch.stack((u / ((W - 1) / 2) - 1, v / ((H - 1) / 2) - 1), dim=3)
        return F.grid_sample(img, grid, mode="bilinear", padding_mode="border", align_corners=True)
    if sampler != "explicit":
        raise ValueError(sampler)
    ix = _unnormalise(u, W)
    iy = _unnormalise(v, H)
    x_w = torch.floor(ix)
    y_n = torch.floor(iy)
    x_e = x_w + 1
    y_s = y_n + 1
    w_nw = ((x_e - ix) * (y_s - iy)).unsqueeze(1)
    w_ne = ((ix - x_w) * (y_s - iy)).unsqueeze(1)
    w_sw = ((x_e - ix) * (iy - y_n)).unsqueeze(1)
    w_se = ((ix - x_w) * (iy - y_n)).unsqueeze(1)
    xi_w = x_w.long().clamp(0, W - 1)
    yi_n = y_n.long().clamp(0, H - 1)
    xi_e = (xi_w + 1).clamp(max=W - 1)      # weight is exactly 0 when this clamps
    yi_s = (yi_n + 1).clamp(max=H - 1)
    flat = img.reshape(B, C, H * W)

    def tap(yi, xi):
        idx = (yi * W + xi).view(B, 1, H * W).expand(B, C, H * W)
        return flat.gather(2, idx).view(B, C, H, W)

    out = tap(yi_n, xi_w) * w_nw + tap(yi_n, xi_e) * w_ne
    out = out + tap(yi_s, xi_w) * w_sw
    out = out + tap(yi_s, xi_e) * w_se
    return out


def occlusion(inv_z, invD1, u, v, sampler="explicit") -> torch.Tensor:
    """True where the warped pixel is out of view or fails the z-buffer test (geo:334-350)."""
    B, _, H, W = inv_z.shape
    d1w = sample_border(invD1, u, v, sampler)
    ok = (inv_z > d1w - OCC_THRES) & (u > 0) & (u < W) & (v > 0) & (v < H)
    return ~ok


def skew(w: torch.Tensor) -> torch.Tensor:
    """(N,3) -> (N,3,3) cross-product matrices (geo:146-161)."""
    o = torch.zeros_like(w[:, 0])
    return torch.stack((o, -w[:, 2], w[:, 1], w[:, 2], o, -w[:, 0], -w[:, 1], w[:, 0], o), 1).view(-1, 3, 3)


def rodrigues(w: torch.Tensor) -> torch.Tensor:
    """so(3) exponential without a small-angle guard, as in geo:163-185."""
    B = w.shape[0]
    theta = w.norm(p=2, dim=1).view(B, 1)
    Kx = skew(w / theta)
    eye = torch.eye(3, dtype=w.dtype, device=w.device).expand(B, 3, 3)
    return eye + Kx * torch.sin(theta).view(B, 1, 1) + Kx.bmm(Kx) * (1 - torch.cos(theta)).view(B, 1, 1)


# --------------------------------------------------------------------------- normal equations
def normal_equations(J: torch.Tensor, r: torch.Tensor, reduction: str = "bmm"):
    """J (B,C,HW,6), r (B,C,H,W) -> A=sum J J^T (B,6,6), b=sum J r (B,6,1) (alg:812-834).

    ``"bmm"`` keeps the reference's per-pixel CxCx6 products followed by a sum over
    pixels; ``"einsum"`` is a single contraction (same value up to summation order).
    """
    B, C, HW, _ = J.shape
    if reduction == "einsum":
        A = torch.einsum("bcpi,bcpj->bij", J, J)
        b = torch.einsum("bcpi,bcp->bi", J, r.reshape(B, C, HW)).unsqueeze(2)
        return A, b
    Jp = J.permute(0, 2, 1, 3).reshape(B * HW, C, 6)
    rp = r.reshape(B, C, HW, 1).permute(0, 2, 1, 3).reshape(B * HW, C, 1)
    A = torch.bmm(Jp.transpose(1, 2), Jp).view(B, HW, 6, 6).sum(dim=1)
    b = torch.bmm(Jp.transpose(1, 2), rp).view(B, HW, 6, 1).sum(dim=1)
    return A, b


def damp(A: torch.Tensor) -> torch.Tensor:
    """H = A + 1e-6 * trace(A) * I (alg:2094-2103)."""
    eye = torch.eye(6, dtype=A.dtype, device=A.device).view(1, 6, 6)
    tr = (A * eye).sum(dim=(1, 2))
    return A + (tr * 1e-6).view(-1, 1, 1) * eye


def gn_update(Hm: torch.Tensor, b: torch.Tensor, R: torch.Tensor, t: torch.Tensor) -> Pose:
    """xi = H^-1 b; inverse-compositional left update (alg:2017-2054, geo:105-123).

    Note the reference hands (R, t, dR, dt) to a function declared as
    (d_R, d_t, R0, t0): the effect is R <- R*dR, t <- R*dt + t.
    """
    B = Hm.shape[0]
    xi = torch.bmm(torch.inverse(Hm), b.reshape(B, 6, 1))
    dR = rodrigues(-xi[:, :3, 0])
    dt = -torch.bmm(dR, xi[:, 3:])
    R_new = R.bmm(dR)
    t_new = R.bmm(dt) + t.reshape(B, 3, 1)
    return R_new, t_new.reshape(B, 3)


# --------------------------------------------------------------------------- U_IC pieces
def uic_residuals(R, t, invD0, invD1, x0, x1, s0, s1, px, py, K,
                  obj_mask0=None, obj_mask1=None, remove_tru_sigma=False, sampler="explicit"):
    """Weighted feature residual and validity mask of one iteration (alg:1960-2015).

    Returns wres, res, sigma (B,C,H,W) and occ (B,1,H,W) bool.
    """
    B, C, H, W = x0.shape
    u, v, inv_z = project(px, py, invD0, R, t, K)
    occ = occlusion(inv_z, invD1, u, v, sampler)
    if obj_mask0 is not None:
        occ = occ | ~obj_mask0.view(B, 1, H, W)
    if obj_mask1 is not None:
        occ = occ | ~(sample_border(obj_mask1.to(x0.dtype), u, v, sampler) > 0)
    f_r = sample_border(x1, u, v, sampler)
    s_r = sample_border(s1, u, v, sampler)
    res = f_r - x0
    sigma = torch.sqrt(s_r.pow(2) + s0.pow(2))
    wres = res / sigma
    if remove_tru_sigma:
        # batch-global extremes; only channel 0 of the comparison is used (alg:1976-1978)
        tru = (s_r == s_r.min()) | (s_r == s_r.max()) | (s0 == s0.min()) | (s0 == s0.max())
        occ = occ | tru[:, 0:1]
    wres = torch.where(occ, torch.full_like(wres, EPS_UIC), wres)
    return wres, res, sigma, occ


def uic_jacobian(res, sigma, s0, gf, gs, Ju, Jv):
    """Per-channel 6-vector Jacobian of the weighted residual (alg:867-887).
    gf, gs: tuples (d/dx, d/dy) of unit Sobel gradients of x0 and sigma0.  Not masked."""
    B, C, H, W = s0.shape
    gx = -gf[0] / sigma - res * (s0 * gs[0] / sigma ** 3)
    gy = -gf[1] / sigma - res * (s0 * gs[1] / sigma ** 3)
    J = gx.reshape(B, C, -1, 1) * Ju.view(B, 1, -1, 6) + gy.reshape(B, C, -1, 1) * Jv.view(B, 1, -1, 6)
    return -J


def vertex_map(depth, px, py):
    """geo:1129-1136"""
    return torch.cat((px, py, torch.ones_like(px)), dim=1) * depth


def normal_map(vertex):
    """cross(Sobel_x V, Sobel_y V) normalised; zero at the batch-global depth extremes
    (alg:2148-2171)."""
    B, _, H, W = vertex.shape
    dx, dy = sobel_unit(vertex, normalise=False)
    n = torch.cross(dx, dy, dim=1)
    n = n / (n.norm(p=2, dim=1, keepdim=True) + 1e-8)
    z = vertex[:, 2:3]
    bad = (z == z.min()) | (z == z.max())
    return torch.where(bad.expand(B, 3, H, W), torch.zeros_like(n), n)


def icp_sigma(z0, n_rot):
    """TUM stereo noise model projected on the rotated normal (alg:975-997)."""
    focal, s_xy, s_disp, baseline = 525.0, 5.5, 0.4, 1.2
    s_lat = z0 / focal * s_xy
    s_z = z0 * z0 * s_disp / (focal * baseline)
    sd = torch.cat((s_lat, s_lat, s_z), dim=1)
    return torch.sqrt((n_rot * sd * sd * n_rot).sum(dim=1, keepdim=True) + 1e-8)


def icp_term(V0, V1, N1, R, t, K, obj_mask0=None, obj_mask1=None, sampler="explicit"):
    """Point-to-plane residual r (B,1,H,W), Jacobian (B,1,HW,6) and mask (alg:916-973)."""
    B, _, H, W = V0.shape
    P = torch.bmm(R, V0.reshape(B, 3, H * W)) + t.reshape(B, 3, 1)
    fx, fy, cx, cy = (K[:, i].view(B, 1) for i in range(4))
    u = (P[:, 0] / P[:, 2]) * fx + cx
    v = (P[:, 1] / P[:, 2]) * fy + cy
    inview = (u > 0) & (u < W - 1) & (v > 0) & (v < H - 1)
    V1r = sample_border(V1, u, v, sampler)
    N1r = sample_border(N1, u, v, sampler)
    diff = P.view(B, 3, H, W) - V1r
    occ = ~inview.view(B, 1, H, W) | (diff.norm(p=2, dim=1, keepdim=True) > ICP_DIST)
    if obj_mask0 is not None:
        occ = occ | ~obj_mask0.view(B, 1, H, W)
    if obj_mask1 is not None:
        occ = occ | ~(sample_border(obj_mask1.to(V0.dtype), u, v, sampler) > 0)
    r = (N1r * diff).sum(dim=1, keepdim=True)
    n_rot = torch.bmm(N1r.reshape(B, 3, -1).transpose(1, 2), R)           # rows n^T R, (B,HW,3)
    J_rot = torch.cross(n_rot, V0.reshape(B, 3, -1).transpose(1, 2), dim=2)  # n^T R [V0]x
    J = torch.cat((J_rot, -n_rot), dim=2).view(B, 1, H * W, 6)
    s = icp_sigma(V0[:, 2:3], n_rot.transpose(1, 2).reshape(B, 3, H, W))
    r = r / (s + 1e-8)
    J = -(J / (s.view(B, 1, H * W, 1) + 1e-8))
    r = torch.where(occ, torch.full_like(r, EPS_UIC), r)
    return r, J, occ


def avg_loss(res_list: Sequence[torch.Tensor], invalid: torch.Tensor) -> torch.Tensor:
    """Per-sample sum of squared valid residuals / number of valid pixels (alg:2119-2137)."""
    B, _, H, W = invalid.shape
    n_valid = H * W - invalid.sum(dim=[2, 3]).squeeze()
    tot = torch.zeros_like(invalid, dtype=res_list[0].dtype)
    for r in res_list:
        tot = tot + (torch.where(invalid, torch.zeros_like(r), r) ** 2).sum(dim=1, keepdim=True)
    return tot.sum(dim=[2, 3], keepdim=True).squeeze() / n_valid


# --------------------------------------------------------------------------- U_IC level
def uic_level(pose: Pose, x0, x1, invD0, invD1, K, s0, s1, *, iters: int = 3,
              remove_tru_sigma: bool = False, combine_icp: bool = False,
              depth0=None, depth1=None, scale_func: Optional[Callable] = None, wPrior=None,
              obj_mask0=None, obj_mask1=None, uncer_prop: bool = False,
              sampler: str = "explicit", reduction: str = "bmm",
              trace: Optional[List[Dict]] = None):
    """One pyramid level of TrustRegionInverseWUncertainty.forward (alg:611-723).

    ``trace`` (a list) receives one dict per iteration with A, b, occ and the pose
    the iteration started from.
    """
    B, C, H, W = x0.shape
    R, t = pose
    t = t.reshape(B, 3)
    px, py = pixel_rays(K, H, W)
    if combine_icp:
        V0 = vertex_map(depth0, px, py)
        V1 = vertex_map(depth1, px, py)
        N1 = normal_map(V1)
    gf = sobel_unit(x0)
    gs = sobel_unit(s0)
    Ju, Jv = warp_rows(invD0, K, px, py)
    w_icp = None
    wres = A = None
    for it in range(iters):
        wres, res, sigma, occ = uic_residuals(R, t, invD0, invD1, x0, x1, s0, s1, px, py, K,
                                              obj_mask0, obj_mask1, remove_tru_sigma, sampler)
        J = uic_jacobian(res, sigma, s0, gf, gs, Ju, Jv)
        A, b = normal_equations(J, wres, reduction)
        rec = {"R": R, "t": t, "occ": occ, "A_feat": A, "b_feat": b}
        if combine_icp:
            r_i, J_i, occ_i = icp_term(V0, V1, N1, R, t, K, obj_mask0, obj_mask1, sampler)
            if w_icp is None:
                w_icp = (torch.ones_like(r_i) * 0.01) if scale_func is None else scale_func(r_i, wres, wPrior)
            A_i, b_i = normal_equations(w_icp.view(B, 1, H * W, 1) * J_i, w_icp * r_i, reduction)
            A = A + A_i
            b = b + b_i
            rec.update(occ_icp=occ_i, A_icp=A_i, b_icp=b_i)
        rec.update(A=A, b=b)
        if trace is not None:
            trace.append(rec)
        R, t = gn_update(damp(A), b, R, t)
    weights = w_icp if combine_icp else torch.ones_like(wres)
    if uncer_prop:
        return (R, t), weights, A
    return (R, t), weights


def uic_residual_loss(pose: Pose, x0, x1, invD0, invD1, K, s0, s1, *, remove_tru_sigma=False,
                      combine_icp=False, depth0=None, depth1=None, scale_func=None, wPrior=None,
                      obj_mask0=None, obj_mask1=None, sampler="explicit"):
    """TrustRegionInverseWUncertainty.forward_residuals (alg:725-786)."""
    B, C, H, W = x0.shape
    R, t = pose
    t = t.reshape(B, 3)
    px, py = pixel_rays(K, H, W)
    wres, _, _, occ = uic_residuals(R, t, invD0, invD1, x0, x1, s0, s1, px, py, K,
                                    obj_mask0, obj_mask1, remove_tru_sigma, sampler)
    if not combine_icp:
        return avg_loss([wres], occ)
    V0 = vertex_map(depth0, px, py)
    V1 = vertex_map(depth1, px, py)
    # the reference omits the object masks in this call (alg:768-769)
    r_i, _, occ_i = icp_term(V0, V1, normal_map(V1), R, t, K, sampler=sampler)
    w = (torch.ones_like(r_i) * 0.01) if scale_func is None else scale_func(r_i, wres, wPrior)
    return avg_loss([wres, w * r_i], occ | occ_i)


# --------------------------------------------------------------------------- IC (TrustRegionBase)
def ic_residual(R, t, invD0, invD1, x0, x1, px, py, K, obj_mask0=None, obj_mask1=None,
                sampler="explicit"):
    """compute_warped_residual (alg:1919-1957): r = x1(u,v) - x0, masked entries = 1e-3."""
    B, C, H, W = x0.shape
    u, v, inv_z = project(px, py, invD0, R, t, K)
    occ = occlusion(inv_z, invD1, u, v, sampler)
    r = sample_border(x1, u, v, sampler) - x0
    if obj_mask0 is not None:
        occ = occ | ~obj_mask0.view(B, 1, H, W)
    if obj_mask1 is not None:
        occ = occ | ~(sample_border(obj_mask1.to(x0.dtype), u, v, sampler) > 0)
    r = torch.where(occ.expand(B, C, H, W), torch.full_like(r, EPS_IC), r)
    return r, occ


def direct_solver(A, Jt, weights, r, pose: Pose, invD0, invD1, x0, x1, K, *,
                  solver: str = "Direct-Nodamping", net: Optional[Callable] = None,
                  samples: int = 10, obj_mask1=None, sampler="explicit",
                  trace: Optional[Dict] = None) -> Pose:
    """DirectSolverNet.forward, inverse direction (alg:1604-1691)."""
    B, C, H, W = x0.shape
    R, t = pose
    b = torch.bmm(Jt, (weights * r).reshape(B, -1, 1))
    if solver == "Direct-Nodamping":
        Hm = damp(A)
    elif solver == "Direct-ResVol":
        px, py = pixel_rays(K, H, W)
        eye = torch.eye(6, dtype=A.dtype, device=A.device).view(1, 6, 6)
        diagA = eye * A
        eps = (diagA.sum(dim=(2, 1)) * 1e-6).view(B, 1, 1) * eye
        lambdas = torch.logspace(-5, 5, samples).to(A)
        volume = []
        for s in range(samples):
            R_s, t_s = gn_update(A + (lambdas[s] * diagA + eps), b, R, t)
            r_s, _ = ic_residual(R_s, t_s, invD0, invD1, x0, x1, px, py, K, obj_mask1=obj_mask1,
                                 sampler=sampler)
            volume.append(torch.bmm(Jt, (weights * r_s).reshape(B, -1, 1)))
        feat = torch.cat((torch.cat(volume, dim=2).reshape(B, -1), A.reshape(B, -1)), dim=1)
        damping = net(feat)
        if trace is not None:
            trace.update(volume=torch.cat(volume, dim=2), damping=damping)
        Hm = A + (eye * damping.view(B, 6, 1) + eps)
    else:
        raise NotImplementedError(solver)
    if trace is not None:
        trace.update(b=b, H=Hm)
    return gn_update(Hm, b, R, t)


def ic_level(pose: Pose, x0, x1, invD0, invD1, K, *, iters: int = 3,
             mest: Optional[Callable] = None, wPrior=None, solver: str = "Direct-Nodamping",
             net: Optional[Callable] = None, samples: int = 10, obj_mask0=None, obj_mask1=None,
             sampler: str = "explicit", trace: Optional[List[Dict]] = None):
    """TrustRegionBase.forward (alg:45-121): J and J^T W J once, then ``iters`` solves."""
    B, C, H, W = x0.shape
    R, t = pose
    t = t.reshape(B, 3)
    px, py = pixel_rays(K, H, W)
    gx, gy = sobel_unit(x0)
    Ju, Jv = warp_rows(invD0, K, px, py)
    J = (gx.reshape(B, C, -1, 1) * Ju.view(B, 1, -1, 6) + gy.reshape(B, C, -1, 1) * Jv.view(B, 1, -1, 6))
    J = J.reshape(B, -1, 6)
    r, occ = ic_residual(R, t, invD0, invD1, x0, x1, px, py, K, obj_mask0, obj_mask1, sampler)
    weights = torch.ones_like(r) if mest is None else mest(r, x0, x1, wPrior)
    A = torch.bmm(J.transpose(1, 2), weights.reshape(B, -1, 1) * J)
    for it in range(iters):
        rec: Dict = {"R": R, "t": t, "occ": occ, "A": A}
        R, t = direct_solver(A, J.transpose(1, 2), weights, r, (R, t), invD0, invD1, x0, x1, K,
                             solver=solver, net=net, samples=samples, obj_mask1=obj_mask1,
                             sampler=sampler, trace=rec)
        if trace is not None:
            trace.append(rec)
        r, occ = ic_residual(R, t, invD0, invD1, x0, x1, px, py, K, obj_mask1=obj_mask1, sampler=sampler)
    return (R, t), weights


def ic_residual_loss(pose: Pose, x0, x1, invD0, invD1, K, *, mest=None, wPrior=None,
                     obj_mask0=None, obj_mask1=None, sampler="explicit"):
    """TrustRegionBase.forward_residuals (alg:123-139)."""
    B, C, H, W = x0.shape
    R, t = pose
    px, py = pixel_rays(K, H, W)
    r, occ = ic_residual(R, t.reshape(B, 3), invD0, invD1, x0, x1, px, py, K, obj_mask0, obj_mask1, sampler)
    w = torch.ones_like(r) if mest is None else mest(r, x0, x1, wPrior)
    return avg_loss([w * r], occ)


# --------------------------------------------------------------------------- coarse-to-fine chain
def track_pyramid(levels: Sequence[Dict], pose: Pose, *, variant: str = "U_IC", iters: int = 3,
                  trace: Optional[List[List[Dict]]] = None, **kw):
    """Coarse-to-fine chain of LeastSquareTracking.forward (LeastSquareTracking.py:345-446).

    ``levels`` is ordered coarse to fine; each entry holds the tensors of that level
    (x0, x1, invD0, invD1, K and, for U_IC, s0, s1 [, depth0, depth1]).  Returns the final
    pose and the list of per-level poses (what train mode stacks, :568-575).
    """
    per_level = []
    for lv in levels:
        tr: Optional[List[Dict]] = [] if trace is not None else None
        if variant == "U_IC":
            out = uic_level(pose, lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"], lv["s0"], lv["s1"],
                            iters=iters, depth0=lv.get("depth0"), depth1=lv.get("depth1"), trace=tr, **kw)
        elif variant == "IC":
            out = ic_level(pose, lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"], iters=iters,
                           trace=tr, **kw)
        else:
            raise NotImplementedError(variant)
        pose = out[0]
        per_level.append(pose)
        if trace is not None:
            trace.append(tr)
    return pose, per_level


def pose_epe_loss(R_est, t_est, R_gt, t_gt, depth, K, invalid=None) -> torch.Tensor:
    """compute_RT_EPE_loss (criterions.py:101-136) on maps that are already at the loss resolution: back-project
    (geometry.py:429-445), transform by the target and by each of the N estimated poses (geometry.py:376-399),
    per-sample mean point distance over the valid pixels (EPE3D_loss, criterions.py:22-46), summed over the poses.
    R_est (B,N,3,3), t_est (B,N,3), depth (B,1,h,w), invalid (B,1,h,w) or None -> (B,)."""
    B, _, H, W = depth.shape
    px, py = pixel_rays(K, H, W)
    xyz = torch.cat((px * depth, py * depth, depth), dim=1).reshape(B, 3, -1)
    target = (t_gt.reshape(B, 3, 1) + torch.bmm(R_gt, xyz)).detach()
    mask = torch.isnan(target).any(dim=1)
    if invalid is not None:
        mask = mask | (invalid.reshape(B, -1) > 0)
    total = torch.zeros((B,), dtype=depth.dtype)
    for n in range(R_est.shape[1]):
        est = t_est[:, n].reshape(B, 3, 1) + torch.bmm(R_est[:, n], xyz)
        epe = torch.norm(target - est, p=2, dim=1)
        per = []
        for b in range(B):
            v = epe[b][~mask[b]]
            per.append(v.mean() if v.numel() else torch.zeros((), dtype=depth.dtype))
        total = total + torch.stack(per)
    return total

