"""Training criterion that follows the solver in every iteration of the reference's train loop, backed by the CUDA
library: ``compute_RT_EPE_loss`` with the signature of ``/root/reference/code/models/criterions.py:101-136``.

The resize of the depth / invalid maps to 60x80 stays ``torch.nn.functional.interpolate`` (plumbing, exactly the
reference's call); the back-projection, the N + 1 rigid transforms, the point distances and the per-sample masked
means -- a Python loop over the batch with boolean indexing in the reference -- are one kernel launch, and the
backward (w.r.t. the estimated poses; the target is a constant as in the reference) another.
"""
from __future__ import annotations

import torch
import torch.nn.functional as func

from . import _lib


class _PoseEpeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, invalid, K, R_gt, t_gt, R_est, t_est):
        L = _lib.lib()
        B, N = int(R_est.shape[0]), int(R_est.shape[1])
        h, w = int(depth.shape[-2]), int(depth.shape[-1])
        dev = depth.device
        t = [x.detach().float().contiguous() for x in (depth, K, R_gt, t_gt, R_est, t_est)]
        inv = invalid.detach().float().contiguous() if invalid is not None else None
        loss = torch.empty((B,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            code = L.dpft_pose_epe_loss(t[0].data_ptr(), inv.data_ptr() if inv is not None else None, t[1].data_ptr(),
                                        t[2].data_ptr(), t[3].data_ptr(), t[4].data_ptr(), t[5].data_ptr(), B, N, h, w,
                                        loss.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(code, "dpft_pose_epe_loss")
        ctx.save_for_backward(*t, *([inv] if inv is not None else []))
        ctx.dims = (B, N, h, w, inv is not None)
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        L = _lib.lib()
        B, N, h, w, has_inv = ctx.dims
        depth, K, R_gt, t_gt, R_est, t_est = ctx.saved_tensors[:6]
        inv = ctx.saved_tensors[6] if has_inv else None
        dev = depth.device
        g_loss = g_loss.float().contiguous()
        g_R, g_t = torch.empty_like(R_est), torch.empty_like(t_est)
        with torch.cuda.device(dev):
            code = L.dpft_pose_epe_loss_backward(depth.data_ptr(), inv.data_ptr() if inv is not None else None,
                                                 K.data_ptr(), R_gt.data_ptr(), t_gt.data_ptr(), R_est.data_ptr(),
                                                 t_est.data_ptr(), B, N, h, w, g_loss.data_ptr(), g_R.data_ptr(),
                                                 g_t.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(code, "dpft_pose_epe_loss_backward")
        return None, None, None, None, None, g_R, g_t


def compute_RT_EPE_loss(R_est, t_est, R_gt, t_gt, depth0, K, invalid=None):
    """Drop-in for reference criterions.py:101-136.  Training call: ``R_est (B,N,3,3)``, ``t_est (B,N,3)`` -> the
    depth and the invalid mask are resized to 60x80 and the loss is summed over the N poses; evaluation call:
    ``R_est (B,3,3)`` at the full resolution.  Returns (B,)."""
    if not depth0.is_cuda:
        raise RuntimeError("compute_RT_EPE_loss needs CUDA tensors: this path has no CPU implementation")
    B, _, H, W = depth0.shape
    if R_est.dim() > 3:
        rH, rW = 60, 80
        rdepth = func.interpolate(depth0, size=(rH, rW), mode='bilinear')
        rinvalid = func.interpolate(invalid.float(), size=(rH, rW), mode='bilinear')   # the reference requires it here
        rK = K.clone()
        rK[:, 0] *= float(rW) / W
        rK[:, 1] *= float(rH) / H
        rK[:, 2] *= float(rW) / W
        rK[:, 3] *= float(rH) / H
        return _PoseEpeFn.apply(rdepth, rinvalid, rK, R_gt.detach(), t_gt.detach(), R_est, t_est.reshape(B, -1, 3))
    return _PoseEpeFn.apply(depth0, invalid, K, R_gt, t_gt, R_est.unsqueeze(1), t_est.reshape(B, 1, 3))
