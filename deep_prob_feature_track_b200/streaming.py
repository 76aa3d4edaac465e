"""Host-resident batches streamed through the solver: the step's inputs live in ONE pinned buffer, two device
buffers alternate, and the upload of step k+1 overlaps the solve of step k (the PCIe link, not the solver,
sets the pace of an end-to-end run: a 120x160 batch of 64 pairs is 222 MB of inputs for 0.7 ms of solve)."""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence, Tuple

import torch


def pack_levels(levels: Sequence[Dict[str, torch.Tensor]], pin: bool):
    """All tensors of a pyramid in ONE flat fp32 buffer (256-byte aligned pieces); returns (flat, layout)."""
    layout, off = [], 0
    for i, lv in enumerate(levels):
        for k, v in lv.items():
            layout.append((i, k, off, tuple(v.shape)))
            off += (v.numel() + 63) // 64 * 64
    flat = torch.empty(off, dtype=torch.float32, pin_memory=pin)
    for (i, k, o, shape) in layout:
        n = 1
        for s in shape:
            n *= s
        flat[o:o + n].view(shape).copy_(levels[i][k])
    return flat, layout


def views(flat: torch.Tensor, layout, n_levels: int) -> List[Dict[str, torch.Tensor]]:
    out: List[Dict[str, torch.Tensor]] = [dict() for _ in range(n_levels)]
    for (i, k, o, shape) in layout:
        n = 1
        for s in shape:
            n *= s
        out[i][k] = flat[o:o + n].view(shape)
    return out


class StreamingSolver:
    """solve(levels_on_device) -> SolveResult is called once per host batch; uploads are double-buffered on a
    side stream, the poses of every step are copied back into a pinned (B,12) buffer."""

    def __init__(self, layout, n_floats: int, n_levels: int, batch: int, device, solve: Callable):
        self.dev = torch.device(device)
        self.flats = [torch.empty(n_floats, dtype=torch.float32, device=self.dev) for _ in range(2)]
        self.views = [views(f, layout, n_levels) for f in self.flats]
        self.out_host = torch.empty((batch, 12), dtype=torch.float32, pin_memory=True)
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.uploaded = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.solve = solve
        for ev in self.consumed:
            ev.record(torch.cuda.current_stream(self.dev))

    def _upload(self, i: int, host_flat: torch.Tensor) -> None:
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[i % 2])     # the solve that last read this buffer is done
            self.flats[i % 2].copy_(host_flat, non_blocking=True)
            self.uploaded[i % 2].record(self.copy_stream)

    def run(self, host_batches: Sequence[torch.Tensor]) -> torch.Tensor:
        """host_batches: pinned flat buffers (pack_levels layout).  Returns the pinned pose rows of the LAST batch
        (valid after the stream is synchronised)."""
        main = torch.cuda.current_stream(self.dev)
        n = len(host_batches)
        self._upload(0, host_batches[0])
        for i in range(n):
            if i + 1 < n:
                self._upload(i + 1, host_batches[i + 1])
            main.wait_event(self.uploaded[i % 2])
            res = self.solve(self.views[i % 2])
            self.consumed[i % 2].record(main)
            self.out_host.copy_(res.pose_hist[-1], non_blocking=True)
        return self.out_host
