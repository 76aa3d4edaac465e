"""Host-side mirror of the reference's solver modules, backed by the CUDA library.

Class names, constructor arguments, ``forward`` signatures, return tuples and attribute names follow
``/root/reference/code/models/algorithms.py`` (TrustRegionInverseWUncertainty :579-997, TrustRegionBase
:23-139) so a ``LeastSquareTracking`` can have its ``tr_update0..3`` children swapped for these
(``patch_tracker`` below) and keep loading its checkpoints.  All arithmetic happens in
``libdpft.so`` (csrc/*.cu) through the C ABI of ``include/dpft.h``; there is no PyTorch fallback.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib

Pose = Tuple[torch.Tensor, torch.Tensor]


# ----------------------------------------------------------------------------- plumbing
def _dev_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: this path has no CPU implementation")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _dev_mask(t: Optional[torch.Tensor], name: str) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    return t.to(torch.uint8).contiguous() if t.dtype != torch.uint8 else t.contiguous()


def pack_pose(pose: Pose) -> torch.Tensor:
    """[R (B,3,3), t (B,3) | (B,3,1)] -> (B,12) rows R|t."""
    R, t = pose
    B = R.shape[0]
    return torch.cat((R.reshape(B, 9), t.reshape(B, 3)), dim=1).float().contiguous()


def unpack_pose(rows: torch.Tensor) -> Pose:
    B = rows.shape[0]
    return rows[:, :9].reshape(B, 3, 3), rows[:, 9:12]


_TRI = [(i, j) for i in range(6) for j in range(i, 6)]
_TRI_OF = [[_TRI.index((min(i, j), max(i, j))) for j in range(6)] for i in range(6)]   # (i,j) -> packed index
_TRI_GATHER: Dict = {}


def _tri_gather(device) -> torch.Tensor:
    g = _TRI_GATHER.get(device)
    if g is None:
        g = _TRI_GATHER[device] = torch.tensor(_TRI_OF, dtype=torch.long, device=device).reshape(36)
    return g


_LAMBDAS: Dict[tuple, torch.Tensor] = {}


def _resvol_lambdas(device, samples: int) -> torch.Tensor:
    """The residual volume's trial dampings, logspace(-5, 5, S) (alg:1664): built once per device (the reference builds
    them on the host in every iteration -- a pageable host-to-device copy, i.e. a stream synchronisation, each time)."""
    key = (device, int(samples))
    lam = _LAMBDAS.get(key)
    if lam is None:
        lam = _LAMBDAS[key] = torch.logspace(-5, 5, int(samples)).to(device=device, dtype=torch.float32)
    return lam


def unpack_system(sys_rows: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(...,27) -> symmetric J^T W J (...,6,6) and J^T W r (...,6,1); one gather, no per-entry launches."""
    A = sys_rows.index_select(-1, _tri_gather(sys_rows.device)).reshape(sys_rows.shape[:-1] + (6, 6))
    return A, sys_rows[..., 21:27].unsqueeze(-1)


class SolveResult:
    """What one call into the library hands back (all device tensors, nothing synchronised)."""
    __slots__ = ("pose_hist", "sys_hist", "aux_hist", "status", "occ", "n_levels", "iters", "launch_ms", "queue_kernel_ms")

    def __init__(self, pose_hist, sys_hist, aux_hist, status, occ, n_levels, iters, launch_ms=None, queue_kernel_ms=None):
        self.pose_hist, self.sys_hist, self.aux_hist = pose_hist, sys_hist, aux_hist
        self.status, self.occ = status, occ
        self.n_levels, self.iters, self.launch_ms = n_levels, iters, launch_ms
        self.queue_kernel_ms = queue_kernel_ms     # timed=True on the work-queue path: event time of each queue kernel alone

    @property
    def pose(self) -> Pose:
        return unpack_pose(self.pose_hist[-1])

    def level_pose(self, level: int) -> Pose:
        """Pose after the ``level``-th solved level (0 = coarsest)."""
        return unpack_pose(self.pose_hist[(level + 1) * self.iters])

    def raise_if_bad(self) -> None:
        """Host check of the device status word (one sync) -- the reference's check_nan asserts
        (algorithms.py:886, :1988) raise AssertionError, so does this."""
        st = int(self.status.item())
        assert not (st & _lib.DPFT_ST_NONFINITE), "non-finite weighted residual / normal equations"
        if st & _lib.DPFT_ST_SINGULAR:
            raise RuntimeError("damped Gauss-Newton system is not positive definite")


def default_queue_levels(shapes: Sequence[Tuple[int, int]], B: int, masks: bool = False) -> int:
    """How many of the finest pyramid levels of a work-queue solve run as work-queue launches (``uic_solve``'s rule when
    ``queue_levels`` is 0).  ``shapes``: (H, W) per level, coarse first; ``B``: pairs of the call; ``masks``: object masks.

    More levels join the queue when a call offers them enough warp rows AND enough pairs to hide a pair's iteration
    boundary (fold, solve, append) behind the tiles of the others.  A level must qualify for the staged routine
    (W >= 60) or its narrow form (W < 44, no object masks); the plain routine on the queue loses to the
    launch-per-iteration kernels.  Measured on one B200, us per call (profiles/r2/r2e_narrow_probe.txt,
    r2e_vga_levels_probe.txt), levels on the queue 1 / 2 / 3 / 4:
      120x160, batches of 64:  2 batches 716 / 765 / 790 / 851,  4: 1205 / 1152 / 1151 / 1204,  8: 2050 / - / 1875 / 1915,
                               12: 2838 / 2724 / 2667 / 2706,  20: - / 4371 / 4203 / 4222
      480x640, one keyframe:   16 frames 1018 / 1060 / 1087 / 1124,  64 frames 3212 / 3126 / 3152 / 3172
    """
    rows = 1776 * 12     # warp rows one wave of workers walks in a 12-row tile

    def offers(i):
        h, w = shapes[i]
        return B * ((w + 29) // 30) * h if (w % 4 == 0 and (w >= 60 or (8 <= w < 44 and not masks))) else 0

    if len(shapes) >= 2 and B >= 64 and offers(-2) >= 2 * rows:
        if len(shapes) >= 3 and B >= 256 and offers(-3) >= rows:
            return 3
        return 2
    return 1


def uic_solve(levels: Sequence[Dict[str, torch.Tensor]], pose: Pose, *, iters: int = 3,
              remove_tru_sigma: bool = False, combine_icp: bool = False, w_icp: float = 0.01,
              want_occ: bool = False, pdl: bool = True, timed: bool = False, fused_sobel: bool = True, single_launch: bool = False, staged_footprint: bool = True,
              shared_keyframe: bool = False, pairwise_extremes: bool = False,
              obj_mask0: Optional[Sequence] = None, obj_mask1: Optional[Sequence] = None,
              queue: Optional[bool] = None, group: int = 0, tile_rows: Optional[Sequence[int]] = None, queue_ctas: int = 0,
              queue_levels: int = 0, icp_weight: Optional[Sequence[Optional[torch.Tensor]]] = None,
              tuning: Optional[Dict[str, int]] = None) -> SolveResult:
    """Coarse-to-fine U_IC solve of a batch of frame pairs on the current CUDA stream.

    ``levels`` is ordered coarse to fine; each dict holds x0, x1, s0, s1 (B,C,h,w), invD0, invD1 (B,1,h,w)
    and K (B,4) already scaled to the level.  One entry == one TrustRegionInverseWUncertainty.forward.

    s0 / s1 may also be (B,1,h,w): the single uncertainty map the reference's encoder produces before repeating
    it to C channels (algorithms.py:1425-1427).  The fused launch-per-iteration kernels then read it once per
    pixel (DPFT_SIGMA_BROADCAST); every other path gets the repeated tensor, so results never depend on it.

    ``group``: the B pairs are B / group independent batches of ``group`` consecutive pairs, each with its own
    batch-global sigma extremes (alg:1976-1979) -- the results of B / group separate calls, from one call.  0 = one
    batch.  ``queue``: the finest level runs as ONE launch whose warps take tiles from a work queue with per-pair
    dependencies (csrc/uic_queue.cu) instead of one launch per iteration; it needs C == 8, fused Sobel, no ICP term
    and no ``want_occ``.  ``None`` (default) picks it when it pays: more than one group, or pairs that nothing
    couples, and at least two waves of tiles at the finest level.  ``queue_levels``: how many of the finest levels take
    that path (0: one, or two / three when the call holds enough pairs and warp rows for the coarser levels to fill the
    queue as well -- levels narrower than 44 columns then run the staged routine's narrow form).  ``tile_rows`` (per level, coarse first), ``queue_ctas`` and ``tuning`` (cta_slots,
    tiling, generic_geometry) are measurement knobs.  ``icp_weight``: with ``combine_icp``, per level a (B,1,h,w) map
    that scales the point-to-plane term pixel by pixel (a learned ScaleNet's output, alg:677-682) instead of ``w_icp``.
    """
    L = _lib.lib()
    n_levels = len(levels)
    x0 = levels[0]["x0"]
    B, C = int(levels[0]["x1"].shape[0]), int(x0.shape[1])
    Bk = 1 if shared_keyframe else B     # batch size of the keyframe-side tensors
    one_sigma = C > 1 and all(int(lv[k].shape[1]) == 1 for lv in levels for k in ("s0", "s1"))
    sigma_broadcast = one_sigma and fused_sobel and not single_launch
    eff_group = group if group > 0 else (1 if pairwise_extremes else B)
    if B % eff_group:
        raise ValueError(f"group ({eff_group}) must divide the batch size ({B})")
    n_groups = B // eff_group
    pairwise_extremes = pairwise_extremes or (eff_group == 1 and B > 1)
    if n_groups > 1 and (combine_icp or want_occ or not fused_sobel or single_launch):
        raise NotImplementedError("sigma-extreme groups are served by the fused launch-per-iteration and work-queue kernels only")
    queue_ok = fused_sobel and not single_launch and not combine_icp and not want_occ and C == 8 and iters >= 1
    if queue is None:
        Hf, Wf = int(levels[-1]["x1"].shape[2]), int(levels[-1]["x1"].shape[3])
        coupled = remove_tru_sigma and n_groups == 1 and B > 1
        queue = (not coupled) and B * ((Wf + 29) // 30) * Hf >= 2 * 1776 * 12
    use_queue = bool(queue) and queue_ok
    if use_queue and queue_levels == 0:
        queue_levels = default_queue_levels([(int(lv["x1"].shape[2]), int(lv["x1"].shape[3])) for lv in levels], B,
                                            obj_mask0 is not None or obj_mask1 is not None)
    SC = 1 if sigma_broadcast else C
    dev = x0.device
    keep = []   # keep converted tensors alive until the launches are queued
    arr = (_lib.DpftLevel * n_levels)()
    occ: List[Optional[torch.Tensor]] = []
    for i, lv in enumerate(levels):
        t = {k: _dev_f32(lv[k], k) for k in ("x0", "x1", "invD0", "invD1", "K")}
        for k in ("s0", "s1"):
            t[k] = _dev_f32(lv[k].expand(-1, C, -1, -1) if (one_sigma and not sigma_broadcast) else lv[k], k)
        H, W = int(t["x0"].shape[2]), int(t["x0"].shape[3])
        for k, nb, nc in (("x0", Bk, C), ("x1", B, C), ("s0", Bk, SC), ("s1", B, SC)):
            if tuple(t[k].shape) != (nb, nc, H, W):
                raise ValueError(f"level {i}: {k} has shape {tuple(t[k].shape)}, expected {(nb, nc, H, W)}")
        for k, nb in (("invD0", Bk), ("invD1", B)):
            if t[k].numel() != nb * H * W:
                raise ValueError(f"level {i}: {k} must be ({nb},1,H,W)")
        if tuple(t["K"].shape) != (B, 4):
            raise ValueError(f"level {i}: K must be (B,4)")
        m0 = _dev_mask(obj_mask0[i] if obj_mask0 is not None else None, "obj_mask0")
        m1 = _dev_mask(obj_mask1[i] if obj_mask1 is not None else None, "obj_mask1")
        o = torch.empty((iters, B, H, W), dtype=torch.uint8, device=dev) if want_occ else None
        occ.append(o)
        keep += [t, m0, m1]
        a = arr[i]
        a.x0, a.x1, a.sigma0, a.sigma1 = (t[k].data_ptr() for k in ("x0", "x1", "s0", "s1"))
        a.invd0, a.invd1, a.K = t["invD0"].data_ptr(), t["invD1"].data_ptr(), t["K"].data_ptr()
        if combine_icp:
            dd = (_dev_f32(lv["depth0"], "depth0"), _dev_f32(lv["depth1"], "depth1"))
            keep.append(dd)
            a.depth0, a.depth1 = dd[0].data_ptr(), dd[1].data_ptr()
        else:
            a.depth0 = a.depth1 = None
        a.obj_mask0 = m0.data_ptr() if m0 is not None else None
        a.obj_mask1 = m1.data_ptr() if m1 is not None else None
        a.occ_out = o.data_ptr() if o is not None else None
        a.H, a.W = H, W
    flags = ((_lib.DPFT_REMOVE_TRU_SIGMA if remove_tru_sigma else 0) | (0 if pdl else _lib.DPFT_NO_PDL)
             | (_lib.DPFT_FUSED_SOBEL if fused_sobel else 0) | (_lib.DPFT_COMBINE_ICP if combine_icp else 0)
             | (0 if single_launch else _lib.DPFT_LAUNCH_PER_ITERATION)
             | (_lib.DPFT_STAGED_FOOTPRINT if staged_footprint else 0)
             | (_lib.DPFT_SHARED_KEYFRAME if shared_keyframe else 0)
             | (_lib.DPFT_PAIRWISE_EXTREMES if pairwise_extremes else 0)
             | (_lib.DPFT_SIGMA_BROADCAST if sigma_broadcast else 0)
             | (_lib.DPFT_QUEUE if use_queue else 0))
    if (shared_keyframe or pairwise_extremes) and (combine_icp or not fused_sobel):
        raise NotImplementedError("shared_keyframe / pairwise_extremes are served by the fused U_IC kernel only")
    n_it = n_levels * iters
    pose_in = pack_pose(pose).to(dev)
    pose_hist = torch.empty((n_it + 1, B, 12), dtype=torch.float32, device=dev)
    sys_hist = torch.empty((max(n_it, 1), B, 27), dtype=torch.float32, device=dev)
    aux_shape = (max(n_it, 1), 4) if n_groups == 1 else (max(n_it, 1), n_groups, 4)
    aux_hist = torch.zeros(aux_shape, dtype=torch.float32, device=dev)
    status = torch.zeros((1,), dtype=torch.int32, device=dev)
    buf = (ctypes.c_float * max(n_it, 1))() if timed else None
    wmaps = None
    if icp_weight is not None:
        if not combine_icp or len(icp_weight) != n_levels:
            raise ValueError("icp_weight needs combine_icp and one entry per level")
        wmaps = [None if w is None else _dev_f32(w, "icp_weight") for w in icp_weight]
        for w, lv in zip(wmaps, levels):
            if w is not None and w.numel() != B * int(lv["x0"].shape[2]) * int(lv["x0"].shape[3]):
                raise ValueError("icp_weight maps must be (B,1,h,w)")
        keep.append(wmaps)
    opt = _lib.DpftUicOptions(group=eff_group, tile_rows=tile_rows, queue_ctas=queue_ctas, queue_levels=queue_levels,
                              icp_weight=None if wmaps is None else [None if w is None else w.data_ptr() for w in wmaps],
                              **(tuning or {}))
    qbuf = (ctypes.c_float * max(1, queue_levels))() if (timed and use_queue) else None
    if timed:   # measurement aid (bench.py): per-iteration device times, synchronises the stream
        opt.launch_ms = ctypes.cast(buf, ctypes.POINTER(ctypes.c_float))
        if qbuf is not None:
            opt.queue_kernel_ms = ctypes.cast(qbuf, ctypes.POINTER(ctypes.c_float))
    ws_bytes = L.dpft_uic_workspace_bytes_ex(arr, n_levels, B, C, iters, flags, ctypes.byref(opt))
    if ws_bytes == 0:
        raise RuntimeError("dpft_uic_workspace_bytes: " + L.dpft_last_error().decode())
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        code = L.dpft_uic_forward_ex(arr, n_levels, B, C, iters, flags, ctypes.c_float(w_icp), pose_in.data_ptr(),
                                     pose_hist.data_ptr(), sys_hist.data_ptr(), aux_hist.data_ptr(), status.data_ptr(),
                                     ws.data_ptr(), ws_bytes, stream, ctypes.byref(opt))
    _lib.check(code, "dpft_uic_forward")
    launch_ms = list(buf) if timed else None
    del keep   # launches are queued on the allocating stream; the caching allocator orders reuse after them
    return SolveResult(pose_hist, sys_hist[:n_it], aux_hist[:n_it], status, occ, n_levels, iters, launch_ms,
                       list(qbuf) if qbuf is not None else None)


def depth_pyramids(depth: torch.Tensor, n_levels: int = 4, with_depth: bool = False):
    """Depth stage of LeastSquareTracking._preprocess (LeastSquareTracking.py:656-661, 668-674): returns the
    inverse-depth pyramid [level 0 (finest), 1, ...] (clamp(1/d,0,10), batch-global min / max zeroed, max-pooled)
    and, with ``with_depth``, the max-pooled depth pyramid the ICP term wants."""
    L = _lib.lib()
    d = _dev_f32(depth, "depth")
    B, H, W = int(d.shape[0]), int(d.shape[-2]), int(d.shape[-1])
    dev = d.device
    inv = [torch.empty((B, 1, H >> l, W >> l), dtype=torch.float32, device=dev) for l in range(n_levels)]
    dpt = [torch.empty_like(t) for t in inv] if with_depth else None
    PtrArr = ctypes.c_void_p * n_levels
    ws = torch.empty((8,), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        code = L.dpft_preprocess_depth(d.data_ptr(), B, H, W, n_levels, PtrArr(*[t.data_ptr() for t in inv]),
                                       PtrArr(*[t.data_ptr() for t in dpt]) if dpt else None, ws.data_ptr(), 8,
                                       torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(code, "dpft_preprocess_depth")
    return (inv, dpt) if with_depth else inv


# ----------------------------------------------------------------------------- autograd
def _level_array(levels, B, C, obj_mask0=None, obj_mask1=None, with_depth=False):
    """ctypes array of dpft_level for already-converted device tensors; returns (array, keep-alive list)."""
    arr = (_lib.DpftLevel * len(levels))()
    keep = []
    for i, lv in enumerate(levels):
        t = {k: _dev_f32(lv[k], k) for k in ("x0", "x1", "s0", "s1", "invD0", "invD1", "K")}
        m0 = _dev_mask(obj_mask0[i] if obj_mask0 is not None else None, "obj_mask0")
        m1 = _dev_mask(obj_mask1[i] if obj_mask1 is not None else None, "obj_mask1")
        keep += [t, m0, m1]
        a = arr[i]
        a.x0, a.x1, a.sigma0, a.sigma1 = (t[k].data_ptr() for k in ("x0", "x1", "s0", "s1"))
        a.invd0, a.invd1, a.K = t["invD0"].data_ptr(), t["invD1"].data_ptr(), t["K"].data_ptr()
        a.obj_mask0 = m0.data_ptr() if m0 is not None else None
        a.obj_mask1 = m1.data_ptr() if m1 is not None else None
        if with_depth:
            dd = (_dev_f32(lv["depth0"], "depth0"), _dev_f32(lv["depth1"], "depth1"))
            keep.append(dd)
            a.depth0, a.depth1 = dd[0].data_ptr(), dd[1].data_ptr()
        a.H, a.W = int(t["x0"].shape[2]), int(t["x0"].shape[3])
    return arr, keep


def uic_residual_loss(level: Dict[str, torch.Tensor], pose: Pose, *, remove_tru_sigma: bool = False,
                      combine_icp: bool = False, w_icp: float = 0.01, obj_mask0=None, obj_mask1=None) -> torch.Tensor:
    """Per-pair average squared weighted residual at ``pose`` (reference forward_residuals, alg:725-786)."""
    L = _lib.lib()
    x0 = level["x0"]
    B, C, dev = int(x0.shape[0]), int(x0.shape[1]), x0.device
    level = dict(level, s0=level["s0"].expand(-1, C, -1, -1), s1=level["s1"].expand(-1, C, -1, -1))
    arr, keep = _level_array([level], B, C, None if obj_mask0 is None else [obj_mask0],
                             None if obj_mask1 is None else [obj_mask1], with_depth=combine_icp)
    flags = (_lib.DPFT_REMOVE_TRU_SIGMA if remove_tru_sigma else 0) | (_lib.DPFT_COMBINE_ICP if combine_icp else 0)
    pose_in = pack_pose(pose).to(dev)
    loss = torch.empty((B,), dtype=torch.float32, device=dev)
    ws_bytes = L.dpft_uic_residual_workspace_bytes(arr, B, C, flags)
    ws = torch.empty((max(ws_bytes, 1),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        code = L.dpft_uic_residual_loss(arr, B, C, flags, ctypes.c_float(w_icp), pose_in.data_ptr(), loss.data_ptr(),
                                        ws.data_ptr(), ws_bytes, torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(code, "dpft_uic_residual_loss")
    del keep
    return loss


def uic_icp_context(level: Dict[str, torch.Tensor], pose: Pose, *, remove_tru_sigma: bool = False, obj_mask0=None,
                    obj_mask1=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """What a learned ScaleNet reads at the first iteration of a level (reference alg:677-680, 1535-1567):
    ``icp_r`` (B,1,H,W), the point-to-plane residual with 1e-6 where its mask is set, and ``feat_norm`` (B,1,H,W) =
    sqrt(sum_c wres_c^2) of the masked uncertainty-weighted feature residual -- written by the residual kernels, the
    (B,C,H,W) residual map is never formed.  ``ScaleNet.compute_rtr`` of them gives exactly its two inputs."""
    L = _lib.lib()
    x0 = level["x0"]
    B, C, H, W, dev = int(x0.shape[0]), int(x0.shape[1]), int(x0.shape[2]), int(x0.shape[3]), x0.device
    level = dict(level, s0=level["s0"].expand(-1, C, -1, -1), s1=level["s1"].expand(-1, C, -1, -1))
    arr, keep = _level_array([level], B, C, None if obj_mask0 is None else [obj_mask0],
                             None if obj_mask1 is None else [obj_mask1], with_depth=True)
    flags = (_lib.DPFT_REMOVE_TRU_SIGMA if remove_tru_sigma else 0) | _lib.DPFT_COMBINE_ICP
    pose_in = pack_pose(pose).to(dev)
    icp_r = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev)
    feat_norm = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev)
    ws_bytes = L.dpft_uic_residual_workspace_bytes(arr, B, C, flags)
    ws = torch.empty((max(ws_bytes, 1),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        code = L.dpft_uic_icp_context(arr, B, C, flags, pose_in.data_ptr(), icp_r.data_ptr(), feat_norm.data_ptr(),
                                      ws.data_ptr(), ws_bytes, torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(code, "dpft_uic_icp_context")
    del keep
    return icp_r, feat_norm


class _UicSolveFn(torch.autograd.Function):
    """Differentiable coarse-to-fine U_IC solve.  Inputs: (cfg, R, t, x0_0, x1_0, s0_0, s1_0, x0_1, ...);
    outputs: per level (R_l, t_l, JtWJ_l).  Backward calls dpft_uic_backward (recompute-based)."""

    @staticmethod
    def forward(ctx, cfg, R, t, *maps):
        n_levels = len(cfg["static"])
        levels = []
        for i, st in enumerate(cfg["static"]):
            x0, x1, s0, s1 = maps[4 * i:4 * i + 4]
            levels.append(dict(x0=x0.detach(), x1=x1.detach(), s0=s0.detach(), s1=s1.detach(), **st))
        res = uic_solve(levels, (R.detach(), t.detach()), iters=cfg["iters"], remove_tru_sigma=cfg["tru"],
                        combine_icp=cfg.get("icp", False), w_icp=cfg.get("w_icp", 0.01),
                        obj_mask0=cfg.get("obj_mask0"), obj_mask1=cfg.get("obj_mask1"))
        if cfg.get("check", True):
            res.raise_if_bad()
        ctx.cfg, ctx.res, ctx.levels = cfg, res, levels
        outs = []
        for l in range(n_levels):
            Rl, tl = res.level_pose(l)
            A, _ = unpack_system(res.sys_hist[(l + 1) * cfg["iters"] - 1])
            outs += [Rl.contiguous(), tl.contiguous(), A]
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        cfg, res, levels = ctx.cfg, ctx.res, ctx.levels
        L = _lib.lib()
        n_levels, iters = len(levels), cfg["iters"]
        x0 = levels[0]["x0"]
        B, C, dev = int(x0.shape[0]), int(x0.shape[1]), x0.device
        n_it = n_levels * iters
        gpose = torch.zeros((n_it + 1, B, 12), dtype=torch.float32, device=dev)
        gA = None
        for l in range(n_levels):
            gR, gt, gAl = gouts[3 * l:3 * l + 3]
            row = gpose[(l + 1) * iters]
            if gR is not None:
                row[:, :9] += gR.reshape(B, 9)
            if gt is not None:
                row[:, 9:] += gt.reshape(B, 3)
            if gAl is not None:
                if gA is None:
                    gA = torch.zeros((n_levels, B, 36), dtype=torch.float32, device=dev)
                gA[l] = gAl.reshape(B, 36)
        icp = bool(cfg.get("icp", False))
        arr, keep = _level_array(levels, B, C, cfg.get("obj_mask0"), cfg.get("obj_mask1"), with_depth=icp)
        garr = (_lib.DpftLevelGrad * n_levels)()
        gmaps = []
        for i, lv in enumerate(levels):
            g = [torch.zeros_like(lv["x0"], dtype=torch.float32, memory_format=torch.contiguous_format) for _ in range(4)]
            gmaps += g
            garr[i].g_x0, garr[i].g_x1, garr[i].g_sigma0, garr[i].g_sigma1 = (x.data_ptr() for x in g)
        flags = (_lib.DPFT_REMOVE_TRU_SIGMA if cfg["tru"] else 0) | (_lib.DPFT_COMBINE_ICP if icp else 0)
        gin = torch.empty((B, 12), dtype=torch.float32, device=dev)
        ws_bytes = L.dpft_uic_backward_workspace_bytes(arr, n_levels, B, C, iters, flags)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            code = L.dpft_uic_backward(arr, garr, n_levels, B, C, iters, flags,
                                       ctypes.c_float(cfg.get("w_icp", 0.01)), res.pose_hist.data_ptr(),
                                       res.sys_hist.data_ptr(), res.aux_hist.data_ptr(), gpose.data_ptr(),
                                       gA.data_ptr() if gA is not None else None, gin.data_ptr(), ws.data_ptr(),
                                       ws_bytes, torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(code, "dpft_uic_backward")
        del keep
        return (None, gin[:, :9].reshape(B, 3, 3), gin[:, 9:].contiguous(), *gmaps)


def uic_track(levels: Sequence[Dict[str, torch.Tensor]], pose: Pose, *, iters: int = 3, remove_tru_sigma: bool = False,
              combine_icp: bool = False, w_icp: float = 0.01, obj_mask0=None, obj_mask1=None, check: bool = True):
    """Differentiable coarse-to-fine solve: returns [(R_l, t_l, JtWJ_l) for every level], coarse first.
    Gradients flow to x0, x1, s0, s1 of every level and to the starting pose (train.py's loss sums over the
    per-level poses, criterions.py:101-136)."""
    static = [dict(invD0=lv["invD0"], invD1=lv["invD1"], K=lv["K"]) for lv in levels]
    if combine_icp:
        for st, lv in zip(static, levels):
            st.update(depth0=lv["depth0"], depth1=lv["depth1"])
    cfg = dict(static=static, iters=iters, tru=remove_tru_sigma, icp=combine_icp, w_icp=w_icp, obj_mask0=obj_mask0,
               obj_mask1=obj_mask1, check=check)
    R, t = pose
    B = R.shape[0]
    maps = []
    for lv in levels:
        C = lv["x0"].shape[1]
        # a single-channel uncertainty map trains as the repeated tensor (autograd sums its gradient over channels)
        maps += [lv["x0"], lv["x1"], lv["s0"].expand(-1, C, -1, -1), lv["s1"].expand(-1, C, -1, -1)]
    outs = _UicSolveFn.apply(cfg, R, t.reshape(B, 3), *maps)
    return [tuple(outs[3 * l:3 * l + 3]) for l in range(len(levels))]


# ----------------------------------------------------------------------------- reference-shaped modules
class TrustRegionInverseWUncertainty(nn.Module):
    """Drop-in for reference algorithms.py:579-997 (track_type 'U_IC').

    Same constructor and ``forward`` signature and return tuple.  ``mEst_func`` and ``solver_func`` are
    accepted and stored under the reference's attribute names (``mEstimator``, ``directSolver``) because
    checkpoints carry their parameters, but -- exactly as in the reference (:642-644, :836) -- never called.
    ``vis_res`` is accepted and ignored (the reference opens OpenCV windows there).
    """

    def __init__(self, max_iter=3, mEst_func=None, solver_func=None, timers=None, uncer_prop=False,
                 combine_icp=False, scale_func=None, remove_tru_sigma=False):
        super().__init__()
        self.max_iterations = max_iter
        self.mEstimator = mEst_func
        self.directSolver = solver_func
        self.timers = timers
        self.uncer_prop = uncer_prop
        self.combine_icp = combine_icp
        self.scale_func = scale_func
        self.remove_tru_sigma = remove_tru_sigma
        self.check_nan = True   # reference asserts after every iteration (host sync); one sync per call here

    def forward(self, pose10, x0, x1, invD0, invD1, K, sigma0, sigma1, wPrior=None, depth0=None, depth1=None,
                vis_res=True, obj_mask0=None, obj_mask1=None):
        assert sigma0 is not None and sigma1 is not None
        learned = self.combine_icp and self._learned_scaler()
        w_icp = 0.01 if learned else self._icp_weight()
        if self.combine_icp:
            assert depth0 is not None and depth1 is not None
        if self.timers: self.timers.tic('trust-region level solve (fused CUDA)')
        lv = dict(x0=x0, x1=x1, s0=sigma0, s1=sigma1, invD0=invD0, invD1=invD1, K=K)
        weights = torch.ones((1, 1, 1, 1), dtype=x0.dtype, device=x0.device).expand(x0.shape)
        needs_grad = torch.is_grad_enabled() and any(t.requires_grad for t in (x0, x1, sigma0, sigma1, pose10[0], pose10[1]))
        if self.combine_icp:
            lv.update(depth0=depth0, depth1=depth1)
            weights = torch.full((1, 1, 1, 1), w_icp, dtype=x0.dtype, device=x0.device).expand(
                x0.shape[0], 1, x0.shape[2], x0.shape[3])
        if learned:
            # A ScaleNet with a CNN (alg:1501-1568): evaluated once, at the level's first iteration, on the two residual
            # maps; its (B,1,H,W) output scales the ICP term pixel by pixel.  The kernels write what it reads
            # (uic_icp_context); the CNN itself stays the reference's module on cuDNN.
            if needs_grad or any(p.requires_grad and torch.is_grad_enabled() for p in self.scale_func.parameters()):
                raise NotImplementedError("training through a learned ICP scaler is not built (no script of the reference uses one)")
            icp_r, feat_norm = uic_icp_context(lv, pose10, remove_tru_sigma=self.remove_tru_sigma, obj_mask0=obj_mask0,
                                               obj_mask1=obj_mask1)
            with torch.no_grad():
                weights = self.scale_func(icp_r, feat_norm, wPrior).float().contiguous()
            res = uic_solve([lv], pose10, iters=self.max_iterations, remove_tru_sigma=self.remove_tru_sigma,
                            combine_icp=True, icp_weight=[weights],
                            obj_mask0=None if obj_mask0 is None else [obj_mask0],
                            obj_mask1=None if obj_mask1 is None else [obj_mask1])
            if self.check_nan:
                res.raise_if_bad()
            if self.timers: self.timers.toc('trust-region level solve (fused CUDA)')
            if self.uncer_prop:
                A, _ = unpack_system(res.sys_hist[-1])
                return res.pose, weights, A
            return res.pose, weights
        if needs_grad:
            # training: same kernels, recorded for autograd (backward = dpft_uic_backward)
            (R, t, A), = uic_track([lv], (pose10[0], pose10[1]), iters=self.max_iterations,
                                   remove_tru_sigma=self.remove_tru_sigma, combine_icp=self.combine_icp, w_icp=w_icp,
                                   obj_mask0=None if obj_mask0 is None else [obj_mask0],
                                   obj_mask1=None if obj_mask1 is None else [obj_mask1], check=self.check_nan)
            if self.timers: self.timers.toc('trust-region level solve (fused CUDA)')
            return ((R, t), weights, A) if self.uncer_prop else ((R, t), weights)
        res = uic_solve([lv], pose10, iters=self.max_iterations, remove_tru_sigma=self.remove_tru_sigma,
                        combine_icp=self.combine_icp, w_icp=w_icp,
                        obj_mask0=None if obj_mask0 is None else [obj_mask0],
                        obj_mask1=None if obj_mask1 is None else [obj_mask1])
        if self.check_nan:
            res.raise_if_bad()
        if self.timers: self.timers.toc('trust-region level solve (fused CUDA)')
        if self.uncer_prop:
            A, _ = unpack_system(res.sys_hist[-1])
            return res.pose, weights, A
        return res.pose, weights


    def _learned_scaler(self) -> bool:
        return self.scale_func is not None and getattr(self.scale_func, "D", -1) > 0

    def _icp_weight(self) -> float:
        """ScaleNet('None') is a constant (ones * scale, alg:1535,1563-1567); that is what every shipped script uses."""
        if not self.combine_icp or self.scale_func is None:
            return 0.01
        return float(getattr(self.scale_func, "scale", 0.01))

    def forward_residuals(self, pose10, x0, x1, invD0, invD1, K, sigma0, sigma1, wPrior=None, depth0=None,
                          depth1=None, vis_res=True, obj_mask0=None, obj_mask1=None):
        assert sigma0 is not None and sigma1 is not None
        lv = dict(x0=x0, x1=x1, s0=sigma0, s1=sigma1, invD0=invD0, invD1=invD1, K=K)
        if self.combine_icp:
            assert depth0 is not None and depth1 is not None
            lv.update(depth0=depth0, depth1=depth1)
        if self.combine_icp and self._learned_scaler():
            raise NotImplementedError("forward_residuals with a learned ICP scaler is not built")
        return uic_residual_loss(lv, pose10, remove_tru_sigma=self.remove_tru_sigma, combine_icp=self.combine_icp,
                                 w_icp=self._icp_weight(), obj_mask0=obj_mask0, obj_mask1=obj_mask1)


# ----------------------------------------------------------------------------- IC tracker (DeepIC baseline)
def _fc(n_in, n_out):
    return nn.Sequential(nn.Linear(n_in, n_out, True), nn.ReLU(inplace=True))


class DirectSolverNet(nn.Module):
    """Mirror of reference algorithms.py:1583-1691: holds the solver type and, for 'Direct-ResVol', the damping
    MLP (same layer layout, so state_dict keys match).  The solve itself runs in TrustRegionBase below, which
    reads ``type`` / ``samples`` / ``net`` from whichever DirectSolverNet (this one or the reference's) it is given
    -- the reference's ``forward`` wants the dense (B,6,C*H*W) Jacobian, which this implementation never forms."""
    SOLVER_NO_DAMPING = 0
    SOLVER_RESIDUAL_VOLUME = 1

    def __init__(self, solver_type, samples=10, direction='inverse'):
        super().__init__()
        self.direction = direction
        if solver_type == 'Direct-Nodamping':
            self.net = None
            self.type = self.SOLVER_NO_DAMPING
        elif solver_type == 'Direct-ResVol':
            self.samples = samples
            self.net = nn.Sequential(_fc(6 * 6 + 6 * samples, 128), _fc(128, 256), _fc(256, 6))
            self.type = self.SOLVER_RESIDUAL_VOLUME
            for m in self.net.modules():
                if isinstance(m, nn.Linear):
                    nn.init.xavier_uniform_(m.weight)
        else:
            raise NotImplementedError()

    def forward(self, JtJ, Jt, weights, R, pose0, invD0, invD1, x0, x1, K, obj_mask1=None):
        """Reference signature (alg:1604): ``JtJ (B,6,6)``, the dense ``Jt (B,6,C*H*W)``, ``weights`` and the residual
        ``R (B,C,H,W)``, ``pose0 = [R, t]`` -> updated pose.  TrustRegionBase above never calls this (it never forms
        ``Jt``); it exists for callers that hold the dense Jacobian already.  The products with ``Jt`` are torch's bmm
        as in the reference; damping, solve, pose update and the re-warped residuals of the residual volume are the
        library's (dpft_ic_update, dpft_ic_residual)."""
        if self.direction != 'inverse':
            raise NotImplementedError("pose updated should be inverse for this tracker")
        lvl = _IcLevel(x0, x1, invD0, invD1, K, None, obj_mask1)
        B = lvl.B
        Jt = Jt.float()
        w = weights.expand(B, lvl.C, lvl.H, lvl.W).float()
        rows = pack_pose(pose0).to(lvl.dev)
        A21 = JtJ.reshape(B, 36).float().index_select(1, torch.tensor([i * 6 + j for i, j in _TRI], device=lvl.dev)).contiguous()
        b0 = torch.bmm(Jt, (w * R).reshape(B, -1, 1)).reshape(B, 6).contiguous()
        if self.type == self.SOLVER_NO_DAMPING:
            return unpack_pose(lvl.update(0, A21, b0, rows)[0])
        S = int(self.samples)
        lambdas = _resvol_lambdas(lvl.dev, S)
        trial = lvl.update(1, A21, b0, rows, lambdas=lambdas)                      # (S,B,12)
        vol = []
        for s in range(S):                                                         # alg:1676-1683
            r_s, _ = lvl.residual(trial[s].contiguous(), first=False)
            vol.append(torch.bmm(Jt, (w * r_s).reshape(B, -1, 1)).reshape(B, 6))
        feat = torch.cat((torch.stack(vol, dim=2).reshape(B, 6 * S), JtJ.reshape(B, 36).float()), dim=1)
        damp = self.net(feat).float().contiguous()
        return unpack_pose(lvl.update(2, A21, b0, rows, damp=damp)[0])


class _IcLevel:
    """One pyramid level of the IC tracker on the device: converted inputs, the unit gradients of x0, and thin
    differentiable wrappers of the dpft_ic_* entry points (each a torch.autograd.Function whose backward is the
    matching dpft_ic_*_backward call, so autograd can chain them around the M-estimator and the damping MLP)."""

    def __init__(self, x0, x1, invD0, invD1, K, obj_mask0=None, obj_mask1=None):
        self.L = _lib.lib()
        self.t = {k: _dev_f32(v, k) for k, v in dict(x0=x0, x1=x1, invD0=invD0, invD1=invD1, K=K).items()}
        self.B, self.C, self.H, self.W = (int(v) for v in self.t["x0"].shape)
        self.dev = self.t["x0"].device
        self.m0, self.m1 = _dev_mask(obj_mask0, "obj_mask0"), _dev_mask(obj_mask1, "obj_mask1")
        self.status = torch.zeros((1,), dtype=torch.int32, device=self.dev)
        self.gx, self.gy = _IcGradFn.apply(self, self.t["x0"])

    def _arr(self, use_mask0):
        arr = (_lib.DpftLevel * 1)()
        a, t = arr[0], self.t
        a.x0, a.x1, a.invd0, a.invd1, a.K = (t[k].data_ptr() for k in ("x0", "x1", "invD0", "invD1", "K"))
        a.obj_mask0 = self.m0.data_ptr() if (use_mask0 and self.m0 is not None) else None
        a.obj_mask1 = self.m1.data_ptr() if self.m1 is not None else None
        a.H, a.W = self.H, self.W
        return arr

    def _call(self, name, *args):
        with torch.cuda.device(self.dev):
            code = getattr(self.L, name)(*args, torch.cuda.current_stream(self.dev).cuda_stream)
        _lib.check(code, name)

    def raise_if_bad(self) -> None:
        """Host check of the level's status word (one sync), as SolveResult.raise_if_bad does for the U_IC path:
        dpft_ic_update factors the damped system by Cholesky, so a system that is not positive definite (negative
        learned weights or damping) would otherwise travel on as NaN poses; the reference's torch.inverse
        (alg:2085-2092) raises on a singular one."""
        st = int(self.status.item())
        assert not (st & _lib.DPFT_ST_NONFINITE), "non-finite weighted residual / normal equations"
        if st & _lib.DPFT_ST_SINGULAR:
            raise RuntimeError("damped Gauss-Newton system of the IC tracker is not positive definite")

    def residual(self, pose_rows, first):
        """(r (B,C,H,W), occ bool (B,1,H,W)); the keyframe object mask only counts on the first call (alg:65-66, 86-87)."""
        r, occ = _IcResidualFn.apply(self, first, pose_rows, self.t["x0"], self.t["x1"])
        return r, occ.bool()

    def context(self, pose_rows, w_prior):
        """(B,4,H,W) input of DeepRobustEstimator('MultiScale2w') at ``pose_rows`` (first evaluation of a level: the
        keyframe object mask counts)."""
        wp = _dev_f32(w_prior, "wPrior")
        ctx = torch.empty((self.B, 4, self.H, self.W), dtype=torch.float32, device=self.dev)
        rows = pose_rows.contiguous()
        with torch.cuda.device(self.dev):
            code = self.L.dpft_ic_context(self._arr(True), self.B, rows.data_ptr(), wp.data_ptr(), int(wp.shape[-2]),
                                          int(wp.shape[-1]), ctx.data_ptr(), None,
                                          torch.cuda.current_stream(self.dev).cuda_stream)
        _lib.check(code, "dpft_ic_context")
        return ctx

    def _w(self, weights):
        if weights is None:
            return None
        return _dev_f32(weights.expand(self.B, self.C, self.H, self.W), "weights")

    def normal_matrix(self, weights):
        return _IcNormalFn.apply(self, self.gx, self.gy, self._w(weights))

    def rhs(self, weights, poses, first=False):
        """(S,B,6) J^T W r at S poses; ``first``: the residual of the level's first warp, which also honours the
        keyframe object mask (alg:63-66 feeds it to the first solve)."""
        return _IcRhsFn.apply(self, first, self.gx, self.gy, self._w(weights), poses, self.t["x0"], self.t["x1"])

    def update(self, mode, A21, rhs, pose_rows, lambdas=None, damp=None):
        return _IcUpdateFn.apply(self, mode, A21, rhs, pose_rows, lambdas, damp)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


class _IcGradFn(torch.autograd.Function):
    """dpft_ic_gradients / dpft_ic_gradients_backward."""

    @staticmethod
    def forward(ctx, lvl, x0):
        gx, gy = torch.empty_like(x0), torch.empty_like(x0)
        lvl._call("dpft_ic_gradients", lvl._arr(False), lvl.B, lvl.C, gx.data_ptr(), gy.data_ptr())
        ctx.lvl = lvl
        return gx, gy

    @staticmethod
    def backward(ctx, g_gx, g_gy):
        lvl = ctx.lvl
        g_gx, g_gy = g_gx.contiguous(), g_gy.contiguous()
        g_x0 = torch.zeros_like(lvl.t["x0"])
        lvl._call("dpft_ic_gradients_backward", lvl._arr(False), lvl.B, lvl.C, g_gx.data_ptr(), g_gy.data_ptr(),
                  g_x0.data_ptr())
        return None, g_x0


class _IcResidualFn(torch.autograd.Function):
    """dpft_ic_residual / dpft_ic_residual_backward."""

    @staticmethod
    def forward(ctx, lvl, first, rows, x0, x1):
        rows = rows.contiguous()
        r = torch.empty_like(x0)
        occ = torch.empty((lvl.B, 1, lvl.H, lvl.W), dtype=torch.uint8, device=lvl.dev)
        lvl._call("dpft_ic_residual", lvl._arr(first), lvl.B, lvl.C, rows.data_ptr(), r.data_ptr(), occ.data_ptr())
        ctx.lvl, ctx.first = lvl, first
        ctx.save_for_backward(rows)
        ctx.mark_non_differentiable(occ)
        return r, occ

    @staticmethod
    def backward(ctx, g_r, _g_occ):
        lvl = ctx.lvl
        rows, = ctx.saved_tensors
        g_r = g_r.contiguous()
        g_x0, g_x1 = torch.zeros_like(lvl.t["x0"]), torch.zeros_like(lvl.t["x1"])
        g_rows = torch.zeros_like(rows)
        lvl._call("dpft_ic_residual_backward", lvl._arr(ctx.first), lvl.B, lvl.C, rows.data_ptr(), g_r.data_ptr(),
                  g_x0.data_ptr(), g_x1.data_ptr(), g_rows.data_ptr())
        return None, None, g_rows, g_x0, g_x1


class _IcNormalFn(torch.autograd.Function):
    """dpft_ic_normal_matrix / dpft_ic_normal_matrix_backward."""

    @staticmethod
    def forward(ctx, lvl, gx, gy, w):
        A21 = torch.empty((lvl.B, 21), dtype=torch.float32, device=lvl.dev)
        lvl._call("dpft_ic_normal_matrix", lvl._arr(False), lvl.B, lvl.C, gx.data_ptr(), gy.data_ptr(), _ptr(w),
                  A21.data_ptr())
        ctx.lvl, ctx.has_w = lvl, w is not None
        ctx.save_for_backward(gx, gy, *([w] if w is not None else []))
        return A21

    @staticmethod
    def backward(ctx, g_A21):
        lvl = ctx.lvl
        gx, gy = ctx.saved_tensors[:2]
        w = ctx.saved_tensors[2] if ctx.has_w else None
        g_A21 = g_A21.contiguous()
        g_gx, g_gy = torch.empty_like(gx), torch.empty_like(gy)
        g_w = torch.empty_like(w) if w is not None else None
        lvl._call("dpft_ic_normal_matrix_backward", lvl._arr(False), lvl.B, lvl.C, gx.data_ptr(), gy.data_ptr(), _ptr(w),
                  g_A21.data_ptr(), g_gx.data_ptr(), g_gy.data_ptr(), _ptr(g_w))
        return None, g_gx, g_gy, g_w


class _IcRhsFn(torch.autograd.Function):
    """dpft_ic_rhs / dpft_ic_rhs_backward (the residual is recomputed inside both)."""

    @staticmethod
    def forward(ctx, lvl, first, gx, gy, w, poses, x0, x1):
        poses = poses.contiguous()
        S = int(poses.shape[0])
        out = torch.empty((S, lvl.B, 6), dtype=torch.float32, device=lvl.dev)
        lvl._call("dpft_ic_rhs", lvl._arr(first), lvl.B, lvl.C, gx.data_ptr(), gy.data_ptr(), _ptr(w), poses.data_ptr(),
                  S, out.data_ptr())
        ctx.lvl, ctx.first, ctx.has_w = lvl, first, w is not None
        ctx.save_for_backward(gx, gy, poses, *([w] if w is not None else []))
        return out

    @staticmethod
    def backward(ctx, g_rhs):
        lvl = ctx.lvl
        gx, gy, poses = ctx.saved_tensors[:3]
        w = ctx.saved_tensors[3] if ctx.has_w else None
        g_rhs = g_rhs.contiguous()
        g_gx, g_gy = torch.zeros_like(gx), torch.zeros_like(gy)
        g_w = torch.zeros_like(w) if w is not None else None
        g_x0, g_x1 = torch.zeros_like(lvl.t["x0"]), torch.zeros_like(lvl.t["x1"])
        g_poses = torch.zeros_like(poses)
        lvl._call("dpft_ic_rhs_backward", lvl._arr(ctx.first), lvl.B, lvl.C, gx.data_ptr(), gy.data_ptr(), _ptr(w),
                  poses.data_ptr(), int(poses.shape[0]), g_rhs.data_ptr(), g_gx.data_ptr(), g_gy.data_ptr(), _ptr(g_w),
                  g_x0.data_ptr(), g_x1.data_ptr(), g_poses.data_ptr())
        return None, None, g_gx, g_gy, g_w, g_poses, g_x0, g_x1


class _IcUpdateFn(torch.autograd.Function):
    """dpft_ic_update / dpft_ic_update_backward."""

    @staticmethod
    def forward(ctx, lvl, mode, A21, rhs, rows, lambdas, damp):
        A21, rhs, rows = A21.contiguous(), rhs.contiguous(), rows.contiguous()
        damp = damp.contiguous() if damp is not None else None
        S = int(lambdas.numel()) if mode == 1 else 1
        out = torch.empty((S, lvl.B, 12), dtype=torch.float32, device=lvl.dev)
        with torch.cuda.device(lvl.dev):
            code = lvl.L.dpft_ic_update(lvl.B, S, mode, A21.data_ptr(), rhs.data_ptr(), _ptr(lambdas), _ptr(damp),
                                        rows.data_ptr(), out.data_ptr(), None, lvl.status.data_ptr(),
                                        torch.cuda.current_stream(lvl.dev).cuda_stream)
        _lib.check(code, "dpft_ic_update")
        ctx.lvl, ctx.mode, ctx.S, ctx.lambdas = lvl, mode, S, lambdas
        ctx.has_damp = damp is not None
        ctx.save_for_backward(A21, rhs, rows, *([damp] if damp is not None else []))
        return out

    @staticmethod
    def backward(ctx, g_out):
        lvl = ctx.lvl
        A21, rhs, rows = ctx.saved_tensors[:3]
        damp = ctx.saved_tensors[3] if ctx.has_damp else None
        g_out = g_out.contiguous()
        g_A21, g_rhs, g_rows = torch.empty_like(A21), torch.empty_like(rhs), torch.empty_like(rows)
        g_damp = torch.empty_like(damp) if damp is not None else None
        with torch.cuda.device(lvl.dev):
            code = lvl.L.dpft_ic_update_backward(lvl.B, ctx.S, ctx.mode, A21.data_ptr(), rhs.data_ptr(), _ptr(ctx.lambdas),
                                                 _ptr(damp), rows.data_ptr(), g_out.data_ptr(), g_A21.data_ptr(),
                                                 g_rhs.data_ptr(), _ptr(g_damp), g_rows.data_ptr(),
                                                 torch.cuda.current_stream(lvl.dev).cuda_stream)
        _lib.check(code, "dpft_ic_update_backward")
        return None, None, g_A21, g_rhs, g_rows, None, g_damp


def _tri_to_full(A21: torch.Tensor) -> torch.Tensor:
    return A21.index_select(-1, _tri_gather(A21.device)).reshape(A21.shape[:-1] + (6, 6))


def _avg_loss(res_list, invalid):
    """compute_avg_loss (alg:2119-2137) on device tensors."""
    B, _, H, W = invalid.shape
    n_valid = H * W - invalid.sum(dim=[2, 3]).squeeze()
    tot = torch.zeros_like(invalid, dtype=torch.float32)
    for r in res_list:
        tot = tot + (torch.where(invalid, torch.zeros_like(r), r) ** 2).sum(dim=1, keepdim=True)
    return tot.sum(dim=[2, 3], keepdim=True).squeeze() / n_valid


class TrustRegionBase(nn.Module):
    """Drop-in for reference algorithms.py:23-139 (track_type 'IC', the DeepIC baseline).  ``mEst_func`` is called
    exactly as the reference calls it (a module taking the residual map, x0, x1 and the weight prior; ``None``
    means constant ones); ``solver_func`` supplies type / samples / damping MLP (see DirectSolverNet)."""

    def __init__(self, max_iter=3, mEst_func=None, solver_func=None, timers=None):
        super().__init__()
        self.max_iterations = max_iter
        self.mEstimator = mEst_func
        self.directSolver = solver_func
        self.timers = timers
        self.check_nan = True   # one read of the level's status word per call (a host sync); False turns it off

    def _weights(self, r, x0, x1, wPrior):
        if self.mEstimator is None:
            return None
        return self.mEstimator(r, x0, x1, wPrior)

    def _fused_context_ok(self, lvl, wPrior, tensors) -> bool:
        """The convolutional M-estimator's input can come straight from the residual kernel (dpft_ic_context) when the
        estimator is the reference's 4-channel one on a one-channel level and nothing needs a gradient."""
        m = self.mEstimator
        if m is None or getattr(m, "D", None) != 4 or getattr(m, "net", None) is None or lvl.C != 1 or wPrior is None:
            return False
        if torch.is_grad_enabled() and (any(t.requires_grad for t in tensors) or any(p.requires_grad for p in m.parameters())):
            return False
        return True

    def forward(self, pose, x0, x1, invD0, invD1, K, wPrior=None, vis_res=False, obj_mask0=None, obj_mask1=None):
        solver = self.directSolver
        if solver is not None and getattr(solver, "direction", "inverse") != "inverse":
            raise NotImplementedError("pose updated should be inverse for this tracker")
        lvl = _IcLevel(x0, x1, invD0, invD1, K, obj_mask0, obj_mask1)
        rows = pack_pose(pose).to(lvl.dev)
        if self.timers: self.timers.tic('compute warping residuals')
        if self._fused_context_ok(lvl, wPrior, (x0, x1, pose[0], pose[1], wPrior)):
            # [ |r|, x0, x1, up(wPrior) ] written once by the residual kernel (alg:1471-1474 builds it from four tensors)
            with torch.no_grad():
                weights = self.mEstimator.net(lvl.context(rows, wPrior))
        else:
            r, occ = lvl.residual(rows, first=True)
            weights = self._weights(r, lvl.t["x0"], lvl.t["x1"], wPrior)
        if self.timers: self.timers.toc('compute warping residuals')
        A21 = lvl.normal_matrix(weights)
        kind = getattr(solver, "type", 0) if solver is not None else 0
        for it in range(self.max_iterations):
            if self.timers: self.timers.tic('solve x=A^{-1}b')
            b0 = lvl.rhs(weights, rows.unsqueeze(0), first=(it == 0))[0]
            if kind == 0:
                rows = lvl.update(0, A21, b0, rows)[0]
            else:
                S = int(solver.samples)
                lambdas = _resvol_lambdas(lvl.dev, S)
                trial = lvl.update(1, A21, b0, rows, lambdas=lambdas)              # (S,B,12)
                vol = lvl.rhs(weights, trial)                                      # (S,B,6)
                feat = torch.cat((vol.permute(1, 2, 0).reshape(lvl.B, 6 * S), _tri_to_full(A21).reshape(lvl.B, 36)), dim=1)
                damp = solver.net(feat).float().contiguous()
                rows = lvl.update(2, A21, b0, rows, damp=damp)[0]
            if self.timers: self.timers.toc('solve x=A^{-1}b')
        if self.check_nan:
            lvl.raise_if_bad()
        if weights is None:
            weights = torch.ones((1, 1, 1, 1), dtype=torch.float32, device=lvl.dev).expand(x0.shape)
        return unpack_pose(rows), weights

    def forward_residuals(self, pose, x0, x1, invD0, invD1, K, wPrior=None, vis_res=False, obj_mask0=None,
                          obj_mask1=None):
        lvl = _IcLevel(x0, x1, invD0, invD1, K, obj_mask0, obj_mask1)
        r, occ = lvl.residual(pack_pose(pose).to(lvl.dev), first=True)
        w = self._weights(r, lvl.t["x0"], lvl.t["x1"], wPrior)
        return _avg_loss([r if w is None else w * r], occ)


class KeyframeTracker:
    """Keyframe-mode tracking as experiments/kf_vo.py does it (TUM_RGBD.py:334-340: every live frame against
    one fixed keyframe), but batched: the keyframe side (x0, sigma0, inverse depth per level) is uploaded once
    and any number of live frames are solved against it in one call.  Every frame keeps the semantics of a
    B = 1 call of the reference (sigma extremes per pair)."""

    def __init__(self, key_levels: Sequence[Dict[str, torch.Tensor]], iters: int = 3, remove_tru_sigma: bool = True):
        self.key = [{k: _dev_f32(lv[k], k) for k in ("x0", "s0", "invD0")} for lv in key_levels]
        for lv in self.key:
            if lv["x0"].shape[0] != 1:
                raise ValueError("a keyframe has batch size 1")
        self.iters, self.tru = iters, remove_tru_sigma

    def track(self, live_levels: Sequence[Dict[str, torch.Tensor]], pose: Pose, **solve_kw) -> SolveResult:
        """live_levels: per level x1, s1 (B,C,h,w), invD1 (B,1,h,w), K (B,4); pose: starting poses of the B frames;
        ``solve_kw``: measurement knobs of ``uic_solve`` (timed, tile_rows, ...)."""
        levels = [dict(kf, x1=lv["x1"], s1=lv["s1"], invD1=lv["invD1"], K=lv["K"]) for kf, lv in zip(self.key, live_levels)]
        return uic_solve(levels, pose, iters=self.iters, remove_tru_sigma=self.tru, shared_keyframe=True,
                         pairwise_extremes=True, **solve_kw)


def _fused_eval_forward(net: nn.Module):
    """``forward`` for a patched U_IC tracker in eval mode: the reference's own ``_preprocess`` (feature encoder, depth
    pyramids, initial pose) and then ALL pyramid levels in ONE solver call -- what ``LeastSquareTracking.forward``
    (LeastSquareTracking.py:345-446) does through four module calls with Python, a status read and a dozen small
    tensor ops between them.  Anything the one call does not serve (training, ICP term, object masks, uncertainty
    propagation, logging / visualisation, timers) goes to the original forward."""
    original = net.forward

    def forward(img0, img1, depth0, depth1, K, init_only=False, logger=None, iteration=0, vis=False, obj_mask0=None,
                obj_mask1=None, index=None):
        trs = [getattr(net, f"tr_update{i}") for i in (3, 2, 1, 0)]
        tr0 = trs[-1]
        fused = (not net.training and not init_only and logger is None and not vis and obj_mask0 is None
                 and obj_mask1 is None and getattr(net, "track_type", None) == "U_IC" and img0.is_cuda
                 and not getattr(net, "vis_feat_uncer", False) and not getattr(net, "timers", None)
                 and all(isinstance(t, TrustRegionInverseWUncertainty) and not t.combine_icp and not t.uncer_prop
                         and t.max_iterations == tr0.max_iterations and t.remove_tru_sigma == tr0.remove_tru_sigma
                         for t in trs))
        if not fused:
            return original(img0, img1, depth0, depth1, K, init_only=init_only, logger=logger, iteration=iteration,
                            vis=vis, obj_mask0=obj_mask0, obj_mask1=obj_mask1, index=index)
        pre = net._preprocess(img0, img1, depth0, depth1, poseI=None, obj_mask0=None, obj_mask1=None)
        x0, x1, sigma0, sigma1, inv_d0, inv_d1, poseI = pre[2], pre[3], pre[4], pre[5], pre[8], pre[9], pre[12]
        levels = [dict(x0=x0[l], x1=x1[l], s0=sigma0[l], s1=sigma1[l], invD0=inv_d0[l], invD1=inv_d1[l],
                       K=K / float(1 << l)) for l in (3, 2, 1, 0)]
        res = uic_solve(levels, poseI, iters=tr0.max_iterations, remove_tru_sigma=tr0.remove_tru_sigma)
        if all(t.check_nan for t in trs):
            res.raise_if_bad()
        return res.pose

    return forward


def patch_tracker(net: nn.Module, fused_forward: bool = False) -> nn.Module:
    """Swap the ``tr_update0..3`` children of a reference ``LeastSquareTracking`` (U_IC or IC) for the CUDA-backed
    modules, keeping their learned sub-modules so ``state_dict`` keys are unchanged.  ``fused_forward=True``
    additionally lets an eval-mode U_IC forward run all four levels in one solver call (same results; see
    ``_fused_eval_forward``)."""
    if fused_forward:
        net = patch_tracker(net)
        net.forward = _fused_eval_forward(net)
        return net
    for i in range(4):
        name = f"tr_update{i}"
        old = getattr(net, name)
        kind = type(old).__name__
        if kind == "TrustRegionInverseWUncertainty":
            new = TrustRegionInverseWUncertainty(old.max_iterations, old.mEstimator, old.directSolver, old.timers,
                                                 old.uncer_prop, old.combine_icp, old.scale_func,
                                                 old.remove_tru_sigma)
        elif kind == "TrustRegionBase":
            new = TrustRegionBase(old.max_iterations, old.mEstimator, old.directSolver, old.timers)
        else:
            raise NotImplementedError(f"{name} is a {kind}: only the U_IC and IC trackers are on this path")
        setattr(net, name, new)
    return net
