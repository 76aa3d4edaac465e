"""Sharding of independent frame pairs over the GPUs of one box.

Pairs (or keyframe sequences) are independent, so every rank takes a contiguous block and there is no
collective on the data path; the only cross-rank traffic is the timing reduction of the benchmark and an
optional gather of the 12 pose floats per pair.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; block sizes differ by at most one."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device="cpu") -> float:
    """Slowest rank's value (device time of a benchmark region); identity without a process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_poses(pose_rows: torch.Tensor, n_total: int) -> List[torch.Tensor]:
    """All ranks' (n_r,12) pose rows on every rank, in pair order (uneven blocks allowed)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [pose_rows]
    world = dist.get_world_size()
    sizes = [shard_range(n_total, world, r) for r in range(world)]
    cap = max(hi - lo for lo, hi in sizes)
    pad = pose_rows.new_zeros((cap, 12))
    pad[: pose_rows.shape[0]] = pose_rows
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return [o[: hi - lo] for o, (lo, hi) in zip(out, sizes)]
