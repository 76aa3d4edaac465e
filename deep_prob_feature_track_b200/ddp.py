"""Data-parallel training step around the solver (BASELINE config 4; reference train.py:117-192 with the
``nn.DataParallel`` of train.py:295-298 replaced by one process per GPU).

The only collective of this code base lives here: the gradients of the feature encoder (7.3 MB in the U_IC
configuration the reference's scripts train) are averaged over the ranks with NCCL.  The solver's own backward
produces gradients of the feature / uncertainty maps; autograd carries them on into the encoder, whose parameters
are the LAST to receive their gradients.  ``FlatBucketReducer`` keeps every gradient as a view into ONE flat buffer
cut into a few buckets (in the order the gradients become ready) and starts the all-reduce of a bucket from the
post-accumulate hook of its last gradient, on NCCL's stream, while autograd is still working on the earlier layers:
the transfer hides behind the tail of the backward.  ``finish()`` waits for the buckets and scales by 1 / world.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class FlatBucketReducer:
    """params: the parameters to synchronise, in registration order (``module.parameters()``).  Gradients become ready
    roughly in REVERSE registration order, so the buckets are cut from the back.  Call ``zero_grad()`` instead of the
    optimizer's (the gradients must stay views of the flat buffer), run backward, then ``finish()``."""

    def __init__(self, params: Iterable[torch.nn.Parameter], n_buckets: int = 4, process_group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no parameters to reduce")
        self.group = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        dev, dt = self.params[0].device, self.params[0].dtype
        order = list(reversed(self.params))                      # the order gradients are expected in
        total = sum(p.numel() for p in order)
        self.flat = torch.zeros(total, dtype=dt, device=dev)
        self.nbytes = total * self.flat.element_size()
        n_buckets = max(1, min(n_buckets, len(order)))
        target = (total + n_buckets - 1) // n_buckets
        self.bucket_of = {}
        self.bounds: List[List[int]] = []                        # [lo, hi) of every bucket in the flat buffer
        self.members: List[int] = []
        off, lo, count = 0, 0, 0
        for p in order:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            self.bucket_of[p] = len(self.bounds)
            off += n
            count += 1
            if off - lo >= target or p is order[-1]:
                self.bounds.append([lo, off])
                self.members.append(count)
                lo, count = off, 0
        self.ready = [0] * len(self.bounds)
        self.handles: List[Optional[object]] = [None] * len(self.bounds)
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        b = self.bucket_of[p]
        self.ready[b] += 1
        if self.ready[b] == self.members[b] and self.world > 1:
            lo, hi = self.bounds[b]
            self.handles[b] = dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def zero_grad(self) -> None:
        self.flat.zero_()
        for p in self.params:                                    # an optimizer may have detached them
            if p.grad is None or p.grad.data_ptr() < self.flat.data_ptr() or \
                    p.grad.data_ptr() >= self.flat.data_ptr() + self.nbytes:
                raise RuntimeError("a gradient left the flat buffer: use this reducer's zero_grad(), not set_to_none")
        self.ready = [0] * len(self.bounds)
        self.handles = [None] * len(self.bounds)

    def finish(self) -> None:
        """Wait for every bucket (buckets whose hook never completed -- unused parameters -- are reduced now) and
        turn the sums into means."""
        if self.world > 1:
            for b, (lo, hi) in enumerate(self.bounds):
                if self.handles[b] is None:
                    self.handles[b] = dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            for h in self.handles:
                h.wait()
            self.flat.mul_(1.0 / self.world)

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()


def broadcast_parameters(module: torch.nn.Module, src: int = 0, process_group=None) -> None:
    """Same weights and buffers on every rank before the first step (what DistributedDataParallel does at construction)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(process_group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=process_group)
