"""Throughput front ends of the U_IC solve for STREAMS of independent batches (what evaluate.py and kf_vo.py loop
over), owned by the package so that callers -- bench.py included -- get the measured throughput from a public call.

A *batch* is what the reference hands to one ``LeastSquareTracking.forward``: its pairs share the batch-global sigma
extremes of ``remove_tru_sigma`` (algorithms.py:1976-1979).  Batches are independent of each other, and a single one
cannot fill a B200: its coarse pyramid levels are latency chains and its finest level is barely one wave of warp
tiles.  So

* ``BatchedSolver`` takes G batches stacked along dim 0 per call (``group`` = batch size keeps the reference's
  semantics per batch; csrc/uic_queue.cu runs the finest level of all of them as one work-queue launch in which the
  pairs drift apart) and issues consecutive calls round-robin on a few CUDA streams, so the coarse levels of one call
  overlap the finest level of another;
* ``HostStreamSolver`` does the same for batches that live in pinned HOST memory: uploads run on a copy stream into a
  ring of device buffers ahead of the solves, poses come back into pinned memory.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import algorithms as A

Pose = Tuple[torch.Tensor, torch.Tensor]


class BatchedSolver:
    """``submit(levels, pose)`` with ``levels`` holding G * batch pairs (coarse level first, tensors as uic_solve takes
    them) enqueues one call on the next stream and returns its SolveResult at once; ``synchronize()`` waits for all.

    batch             pairs per batch (the reference's batch size; the sigma-extreme group)
    streams           CUDA streams the calls rotate over (2 is enough once a call holds several batches)
    graphs            replay every call from a CUDA graph: the ~20 launches of a call (and the torch allocations around
                      them) are captured the first time a set of input buffers is seen and replayed
                      afterwards, which removes the host's submission time from the critical path of short runs.  The
                      caller keeps ownership of the input buffers and refills them in place; the SolveResult of a set of
                      buffers is the SAME object on every call (its tensors are overwritten by the replay).
    solve_kw          passed on to uic_solve (tile_rows, queue, ...)
    """

    def __init__(self, batch: int, *, iters: int = 3, remove_tru_sigma: bool = True, streams: int = 2,
                 device: Optional[torch.device] = None, graphs: bool = False, **solve_kw):
        self.batch, self.iters, self.tru = int(batch), iters, remove_tru_sigma
        self.dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.streams = [torch.cuda.Stream(device=self.dev) for _ in range(max(1, streams))]
        self.solve_kw = solve_kw
        self.graphs = bool(graphs)
        self._graphs: Dict[tuple, Optional[tuple]] = {}
        self.replays = 0
        self._next = 0
        self.calls = 0

    def _solve(self, levels, pose) -> A.SolveResult:
        return A.uic_solve(levels, pose, iters=self.iters, remove_tru_sigma=self.tru, group=self.batch, **self.solve_kw)

    def _captured(self, st: torch.cuda.Stream, levels, pose):
        """[graph, result, buffers, event of the last replay] of these buffers, capturing on first sight; None when
        capture is not possible.  A graph owns its outputs and workspace, so it replays on any stream but never twice
        at once: a replay waits for the previous one of the same graph."""
        key = (tuple(int(v.data_ptr()) for lv in levels for v in lv.values()),
               tuple(tuple(v.shape) for lv in levels for v in lv.values()), int(pose[0].data_ptr()), int(pose[1].data_ptr()))
        if key in self._graphs:
            return self._graphs[key]
        ent = None
        try:
            with torch.cuda.stream(st):
                self._solve(levels, pose)            # once eagerly: function attributes, allocator, lazy module loads
            st.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                res = self._solve(levels, pose)
            ent = [g, res, (levels, pose), None]     # the graph reads these buffers: keep them alive with it
        except Exception:                            # e.g. a knob that synchronises (timed=True): stay eager
            torch.cuda.synchronize(self.dev)
            ent = None
        self._graphs[key] = ent
        return ent

    def submit(self, levels: Sequence[Dict[str, torch.Tensor]], pose: Pose, after: Optional[torch.cuda.Event] = None) -> A.SolveResult:
        n = int(levels[0]["x1"].shape[0])
        if n % self.batch:
            raise ValueError(f"{n} pairs are not a whole number of batches of {self.batch}")
        st = self.streams[self._next % len(self.streams)]
        self._next += 1
        self.calls += 1
        if after is not None:
            st.wait_event(after)
        if self.graphs:
            ent = self._captured(st, levels, pose)
            if ent is not None:
                if ent[3] is not None:
                    st.wait_event(ent[3])
                with torch.cuda.stream(st):
                    ent[0].replay()
                ent[3] = torch.cuda.Event()
                ent[3].record(st)
                self.replays += 1
                return ent[1]
        with torch.cuda.stream(st):
            return self._solve(levels, pose)

    def wait_for(self, event: torch.cuda.Event) -> None:
        """Every stream waits for ``event`` (e.g. the start mark of a timed region) before its next call."""
        for st in self.streams:
            st.wait_event(event)

    def join(self, stream: Optional[torch.cuda.Stream] = None) -> None:
        """``stream`` (default: the current one) waits for everything submitted so far (no host sync)."""
        stream = stream or torch.cuda.current_stream(self.dev)
        for st in self.streams:
            ev = torch.cuda.Event()
            ev.record(st)
            stream.wait_event(ev)

    def synchronize(self) -> None:
        for st in self.streams:
            st.synchronize()


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


class HostStreamSolver:
    """Batches in pinned host memory (one flat buffer each, ``pack_levels`` layout) -> poses in pinned host memory.

    ``depth`` device buffers form a ring: the copy stream uploads step k + depth - 1 while step k is solved, every
    step's inputs cross the link exactly once and its (n,12) pose rows are read back.  ``solve(levels_on_device)``
    is any callable returning a SolveResult (e.g. ``lambda lv: uic_solve(lv, pose0, ...)``).
    """

    def __init__(self, layout, n_floats: int, n_levels: int, n_pairs: int, device, solve: Callable, depth: int = 3):
        self.dev = torch.device(device)
        self.depth = max(2, depth)
        self.flats = [torch.empty(n_floats, dtype=torch.float32, device=self.dev) for _ in range(self.depth)]
        self.views = [views(f, layout, n_levels) for f in self.flats]
        self.out_host = torch.empty((n_pairs, 12), dtype=torch.float32, pin_memory=True)
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.uploaded = [torch.cuda.Event() for _ in range(self.depth)]
        self.consumed = [torch.cuda.Event() for _ in range(self.depth)]
        self.solve = solve
        self.h2d_bytes_per_step = n_floats * 4
        self.d2h_bytes_per_step = n_pairs * 12 * 4
        for ev in self.consumed:
            ev.record(torch.cuda.current_stream(self.dev))

    def _upload(self, i: int, host_flat: torch.Tensor) -> None:
        s = i % self.depth
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[s])     # the solve that last read this buffer is done
            self.flats[s].copy_(host_flat, non_blocking=True)
            self.uploaded[s].record(self.copy_stream)

    def run(self, host_batches: Sequence[torch.Tensor]) -> torch.Tensor:
        """Returns the pinned pose rows of the LAST batch (valid after the stream is synchronised)."""
        main = torch.cuda.current_stream(self.dev)
        n = len(host_batches)
        ahead = self.depth - 1
        for j in range(min(ahead, n)):
            self._upload(j, host_batches[j])
        for i in range(n):
            if i + ahead < n:
                self._upload(i + ahead, host_batches[i + ahead])
            s = i % self.depth
            main.wait_event(self.uploaded[s])
            res = self.solve(self.views[s])
            self.consumed[s].record(main)
            self.out_host.copy_(res.pose_hist[-1], non_blocking=True)
        return self.out_host


def numa_cpus_of_gpu(index: int) -> Optional[List[int]]:
    """CPUs of the NUMA node the GPU hangs off (sysfs), or None when the box does not say."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:          # 00000000:1B:00.0 -> 0000:1b:00.0
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus: List[int] = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
        return cpus or None
    except Exception:
        return None


def bind_to_gpu_numa_node(index: int) -> Optional[int]:
    """Pin this process to the CPUs next to GPU ``index`` so that pinned buffers it allocates afterwards (first touch)
    and its copy-engine traffic stay on that node; returns the number of CPUs, None if nothing was done."""
    import os
    cpus = numa_cpus_of_gpu(index)
    if not cpus:
        return None
    try:
        allowed = os.sched_getaffinity(0)
        want = allowed & set(cpus)
        if want:
            os.sched_setaffinity(0, want)
            return len(want)
    except Exception:
        pass
    return None
