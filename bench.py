#!/usr/bin/env python
"""Benchmark of the trust-region inverse-compositional solver path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload tum|vga]

A *step* is one coarse-to-fine solve (4 pyramid levels x 3 Gauss-Newton iterations, U_IC with
--remove_tru_sigma as in every script the reference ships) of one batch of synthetic frame pairs.
The metric is frame-pair GN solves per second; rank 0 prints ONE JSON line.

  value      whole-job throughput with the inputs already in HBM (CUDA events, max over ranks); independent
             batches are issued round-robin on --streams CUDA streams (default 8) so the latency-bound coarse
             levels of one batch overlap the finest level of another; --streams 1 gives the latency of one solve
  e2e        same call made with HOST (pinned) buffers: host->device copy of the step's inputs and
             device->host read of the poses inside the timed region
  roofline   dominant kernel = the finest-level Gauss-Newton launch; algorithmic bytes per launch
             (4C+2)*4*H*W*B  /  its average device time (events around every launch, separate pass)
  cpu_baseline  the CPU port of the reference (oracle/) on this box's host cores, bounded sample

N > 1 is launched by torchrun (one rank per GPU); frame pairs are independent, so every rank solves
its own batch and there is no collective on the data path (weak scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[1]: TUM-shape 120x160, batch 64, C=8 features, 4 levels x 3 iterations
    "tum": dict(name="tum120x160_b64_c8_uic_4lvl_x3it", B=64, C=8, H=120, W=160),
    # BASELINE.json configs[2] per-GPU shard: 480x640, 128 pairs per GPU, processed 16 at a time
    "vga": dict(name="vga480x640_b16_c8_uic_4lvl_x3it", B=16, C=8, H=480, W=640),
}
N_LEVELS, ITERS = 4, 3
N_SETS = 4           # distinct input sets rotated between steps so no step finds its inputs in L2


def algorithmic_bytes(B, C, H, W, levels=N_LEVELS, iters=ITERS):
    """every iteration reads x0, x1, sigma0, sigma1 (C channels) and both inverse depths."""
    per_level = [(4 * C + 2) * 4 * (H >> l) * (W >> l) * B for l in range(levels)]
    return iters * sum(per_level), per_level[0]


class ClockSampler:
    """SM clock / throttle reasons sampled through NVML while the GPU is busy (same source as nvidia-smi)."""

    def __init__(self, index: int, period: float = 0.01):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:   # pragma: no cover - NVML missing
            self.nv = None
        self.period = period

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_sample(levels, R0, t0, n_pairs, min_seconds=10.0, max_reps=500):
    """Time the oracle port (reference op chain: grid_sample + permute/bmm/sum) on the host cores."""
    from oracle import ic_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sub = [{k: v[:n_pairs].contiguous() for k, v in lv.items()} for lv in levels]
    pose = (R0[:n_pairs].contiguous(), t0[:n_pairs].contiguous())

    def once():
        with torch.no_grad():
            return O.track_pyramid(sub, pose, iters=ITERS, remove_tru_sigma=True, sampler="grid_sample",
                                   reduction="bmm")

    small = [{k: v[:2].contiguous() for k, v in lv.items()} for lv in levels]
    with torch.no_grad():
        O.track_pyramid(small, (R0[:2], t0[:2]), iters=ITERS, remove_tru_sigma=True, sampler="grid_sample",
                        reduction="bmm")
    times = []
    t_begin = time.perf_counter()
    while len(times) < max_reps and (not times or time.perf_counter() - t_begin < min_seconds):
        t = time.perf_counter()
        once()
        times.append(time.perf_counter() - t)
    return n_pairs / min(times), threads, times


def train_step_leg(A, dev_sets, pose0, B, steps, dev):
    """Solver part of a training step (BASELINE config 4): forward + backward through all 12 iterations, loss =
    a linear functional of every level's pose (what criterions.py:101-136 feeds back)."""
    import torch
    leaves = [[{k: (v.clone().requires_grad_(True) if k in ("x0", "x1", "s0", "s1") else v) for k, v in lv.items()}
               for lv in s] for s in dev_sets[:2]]
    R = pose0[0].clone().requires_grad_(True)
    t = pose0[1].clone().requires_grad_(True)

    def step(i):
        for lv in leaves[i % 2]:
            for k in ("x0", "x1", "s0", "s1"):
                lv[k].grad = None
        outs = A.uic_track(leaves[i % 2], (R, t), iters=ITERS, remove_tru_sigma=True, check=False)
        loss = sum(Rl.sum() + tl.sum() for Rl, tl, _ in outs)
        loss.backward()

    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"what": "solver forward+backward (autograd through 4 levels x 3 iterations), per GPU", "ms_per_step": ms,
            "pairs_per_s": B / (ms * 1e-3), "steps": steps}


def vga_leg(A, rank, dev, args):
    """BASELINE config 3 shape: 480x640 pairs, 16 per call (a 128-pair shard is 8 such calls)."""
    import torch
    from deep_prob_feature_track_b200.synthetic import make_frame_pairs
    wl = WORKLOADS["vga"]
    B, C, H, W = wl["B"], wl["C"], wl["H"], wl["W"]
    data = make_frame_pairs(B, C, H, W, seed=99 + rank, n_levels=N_LEVELS)
    sets = []
    for s in range(2):
        sets.append([{k: torch.roll(v, s, 0).to(dev) for k, v in lv.items()} for lv in data["levels"]])
    pose0 = (data["R0"].to(dev), data["t0"].to(dev))

    def solve(i, **kw):
        return A.uic_solve(sets[i % 2], pose0, iters=ITERS, remove_tru_sigma=True, pdl=not args.no_pdl,
                           fused_sobel=args.fused_sobel, single_launch=args.single_launch, staged_footprint=not args.no_staged, **kw)

    for i in range(3):
        solve(i)
    torch.cuda.synchronize()
    n = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        solve(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    r = solve(0, timed=True)
    bytes_step, bytes_lvl0 = algorithmic_bytes(B, C, H, W)
    lvl0 = sum(r.launch_ms[-ITERS:]) / ITERS
    out = {"workload": wl["name"], "ms_per_step": ms, "pairs_per_s": B / (ms * 1e-3),
           "lvl0_launch_ms": lvl0, "lvl0_algorithmic_GBps": bytes_lvl0 / (lvl0 * 1e-3) / 1e9,
           "step_algorithmic_GBps": bytes_step / (ms * 1e-3) / 1e9}
    # keyframe mode (kf_vo.py --vo_type keyframe): the same live frames against ONE keyframe uploaded once
    key = [{k: lv[k][:1].contiguous() for k in ("x0", "s0", "invD0")} for lv in sets[0]]
    tracker = A.KeyframeTracker(key, iters=ITERS, remove_tru_sigma=True)
    live = [[{k: lv[k] for k in ("x1", "s1", "invD1", "K")} for lv in s] for s in sets]
    for i in range(3):
        tracker.track(live[i % 2], pose0)
    torch.cuda.synchronize()
    e0.record()
    for i in range(n):
        tracker.track(live[i % 2], pose0)
    e1.record()
    torch.cuda.synchronize()
    ms_kf = e0.elapsed_time(e1) / n
    out["keyframe_mode"] = {"what": f"{B} live frames per call against one resident keyframe, B=1 semantics per frame",
                            "ms_per_step": ms_kf, "pairs_per_s": B / (ms_kf * 1e-3)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--workload", default="tum", choices=tuple(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pdl", action="store_true")
    ap.add_argument("--materialised", action="store_true",
                    help="materialise the unit Sobel gradients once per level instead of the fused sliding-window kernel")
    ap.add_argument("--single-launch", action="store_true",
                    help="all levels and iterations in ONE cooperative launch instead of one launch per iteration")
    ap.add_argument("--no-staged", action="store_true", help="plain fused kernel instead of the staged-footprint one (DPFT_STAGED_FOOTPRINT off)")
    ap.add_argument("--streams", type=int, default=8, help="CUDA streams the timed steps are spread over")
    ap.add_argument("--no-extras", action="store_true", help="skip the training-step and 480x640 side measurements")
    ap.add_argument("--no-graphs", action="store_true",
                    help="issue every launch from the host instead of replaying one CUDA graph per solve")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    args.fused_sobel = not args.materialised

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]
    B, C, H, W = wl["B"], wl["C"], wl["H"], wl["W"]
    bytes_step, bytes_lvl0 = algorithmic_bytes(B, C, H, W)

    from deep_prob_feature_track_b200.synthetic import make_frame_pairs

    # ------------------------------------------------------------------ reference arm: CPU port, rank 0 only
    if args.impl == "reference":
        if rank != 0:
            return
        n_cpu = min(B, 16 if args.workload == "tum" else 1)
        data = make_frame_pairs(n_cpu, C, H, W, seed=1234, n_levels=N_LEVELS)
        from oracle import ic_oracle as O
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)

        def step():
            with torch.no_grad():
                O.track_pyramid(data["levels"], (data["R0"], data["t0"]), iters=ITERS, remove_tru_sigma=True,
                                sampler="grid_sample", reduction="bmm")

        for _ in range(args.warmup):
            step()
        t = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = time.perf_counter() - t
        v = n_cpu * args.steps / dt
        sample = f"{n_cpu} of the {B} pairs of a step, {args.steps} steps after {args.warmup} warm-up"
        print(json.dumps({
            "impl": "reference", "metric": "frame-pair GN solves/sec", "value": v, "unit": "pairs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "levels": N_LEVELS, "iters_per_level": ITERS, "variant": "U_IC",
                       "remove_tru_sigma": True},
            "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }))
        return

    # ------------------------------------------------------------------ our arm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback exists)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from deep_prob_feature_track_b200 import algorithms as A
    from deep_prob_feature_track_b200.sharding import max_over_ranks
    from deep_prob_feature_track_b200.streaming import StreamingSolver, pack_levels, views

    data = make_frame_pairs(B, C, H, W, seed=1234 + rank, n_levels=N_LEVELS)
    host_flat, layout = pack_levels(data["levels"], pin=True)
    dev_sets = []
    for s in range(N_SETS):
        # distinct allocations (distinct addresses) with the batch rolled by s: L2 can not serve step k+1 from step k
        rolled = [{k: torch.roll(v, s, 0) for k, v in lv.items()} for lv in data["levels"]]
        f, _ = pack_levels(rolled, pin=False)
        dev_sets.append(views(f.to(dev), layout, N_LEVELS))
    pose0 = (data["R0"].to(dev), data["t0"].to(dev))
    set_bytes = host_flat.numel() * 4

    def solve(levels, **kw):
        return A.uic_solve(levels, pose0, iters=ITERS, remove_tru_sigma=True, pdl=not args.no_pdl, fused_sobel=args.fused_sobel,
                           single_launch=args.single_launch, staged_footprint=not args.no_staged, **kw)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # init + sigma0 min/max per level, then either ONE cooperative launch for all 12 iterations (+ its init) or
    # one launch per iteration (+ 2 Sobel launches per level when gradients are materialised)
    single = args.fused_sobel and args.single_launch
    launches_per_step = 1 + N_LEVELS + (2 if single else N_LEVELS * ITERS + (0 if args.fused_sobel else 2 * N_LEVELS))

    with ClockSampler(local_rank) as clocks:
        # ---- value: inputs resident in HBM.  Batches are independent, so consecutive steps are issued round-robin
        # on `--streams` CUDA streams: the latency-bound coarse levels of one batch overlap the bandwidth-heavy
        # finest level of another.  Timed with events on the launching (default) stream, which every worker
        # stream waits for at the start and which waits for every worker stream at the end.
        main_stream = torch.cuda.current_stream(dev)
        workers = [torch.cuda.Stream(device=dev) for _ in range(max(1, args.streams))]

        # One solve is 17 launches (~0.18 ms of host time) for ~0.3 ms of device time; with several ranks on one host
        # the issue rate, not the GPU, set the pace (4 GPUs: 0.56 ms per step).  So a solve is captured ONCE per
        # (worker stream, input set) into a CUDA graph -- same C-ABI call, same launches, PDL edges included -- and the
        # timed steps replay the graphs.  --no-graphs restores host-issued launches; a failed capture falls back to them.
        graphs = {}
        if not args.no_graphs and not args.single_launch:
            try:
                combos = sorted({(i % len(workers), i % N_SETS) for i in range(len(workers) * N_SETS)})
                for w, k in combos:
                    with torch.cuda.stream(workers[w]):
                        solve(dev_sets[k])                      # scratch pools, function attributes
                torch.cuda.synchronize()
                for w, k in combos:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=workers[w]):
                        out = solve(dev_sets[k])
                    graphs[(w, k)] = (g, out)
                torch.cuda.synchronize()
            except Exception as exc:   # pragma: no cover - depends on the driver
                print(f"[bench] CUDA graph capture failed ({exc!r}); issuing launches from the host", file=sys.stderr)
                graphs = {}
                torch.cuda.synchronize()

        def run_steps(n):
            last = None
            for i in range(n):
                w, k = i % len(workers), i % N_SETS
                with torch.cuda.stream(workers[w]):
                    if graphs:
                        graphs[(w, k)][0].replay()
                        last = graphs[(w, k)][1]
                    else:
                        last = solve(dev_sets[k])
            return last

        def timed(n):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(main_stream)
            for w in workers:
                w.wait_event(e0)
            out = run_steps(n)
            for w in workers:
                done = torch.cuda.Event()
                done.record(w)
                main_stream.wait_event(done)
            e1.record(main_stream)
            return out, e0, e1

        use_graphs = bool(graphs)
        res, _, _ = timed(args.warmup)
        res.raise_if_bad()
        barrier()
        res, e0, e1 = timed(args.steps)
        barrier()
        ms = e0.elapsed_time(e1)
        res.raise_if_bad()
        ms = max_over_ranks(ms, dev)

        # ---- roofline: device time of every GN launch (events around each launch; separate pass)
        lvl0 = []
        per_launch = None
        for i in range(max(3, min(args.steps, 20))):
            r = solve(dev_sets[i % N_SETS], timed=True)
            lvl0 += r.launch_ms[-ITERS:]
            per_launch = r.launch_ms if per_launch is None else [a + b for a, b in zip(per_launch, r.launch_ms)]
        n_timed = max(3, min(args.steps, 20))
        per_launch = [x / n_timed for x in per_launch]
        lvl0_ms = sum(lvl0) / len(lvl0)

        # ---- e2e: host buffers in, poses out, every step.  Two device buffers and a copy stream: the upload of
        # step k+1 runs while step k is solved (the upload is ~5x the solve, so the link sets the pace); every
        # step still uploads all of its inputs from pinned memory and reads its poses back.
        streamer = StreamingSolver(layout, host_flat.numel(), N_LEVELS, B, dev, solve)

        def e2e_run(n):
            streamer.run([host_flat] * n)

        n_e2e = max(3, min(args.steps, 30))
        e2e_run(3)
        barrier()
        e0.record()
        e2e_run(n_e2e)
        e1.record()
        barrier()
        ms_e2e = e0.elapsed_time(e1)
        ms_e2e = max_over_ranks(ms_e2e, dev)

        # ---- side measurements (not the headline): training step of the same workload, 480x640 pairs
        extras = {}
        if not args.no_extras:
            # the same workload with the uncertainty passed as the ONE map per frame the reference's encoder emits
            # (it repeats it to C channels, alg:1425-1427; the headline above moves and reads the repeated tensors)
            one_sets = [[dict(lv, s0=lv["s0"][:, :1].contiguous(), s1=lv["s1"][:, :1].contiguous()) for lv in st]
                        for st in dev_sets]

            def run_one(n):
                for i in range(n):
                    with torch.cuda.stream(workers[i % len(workers)]):
                        solve(one_sets[i % N_SETS])

            e0.record(main_stream)
            for w in workers:
                w.wait_event(e0)
            run_one(args.warmup)
            torch.cuda.synchronize()
            e0.record(main_stream)
            for w in workers:
                w.wait_event(e0)
            run_one(args.steps)
            for w in workers:
                done = torch.cuda.Event()
                done.record(w)
                main_stream.wait_event(done)
            e1.record(main_stream)
            torch.cuda.synchronize()
            ms_one = e0.elapsed_time(e1) / args.steps
            r = solve(one_sets[0], timed=True)
            lvl0_one = sum(r.launch_ms[-ITERS:]) / ITERS
            bytes_lvl0_one = (2 * C + 4) * 4 * H * W * B
            extras["single_sigma_map"] = {
                "what": "sigma0 / sigma1 given as (B,1,h,w) instead of repeated to C channels (DPFT_SIGMA_BROADCAST); "
                        "same results", "ms_per_step": ms_one, "pairs_per_s": B / (ms_one * 1e-3),
                "lvl0_launch_ms": lvl0_one, "lvl0_bytes_read_per_launch": bytes_lvl0_one,
                "lvl0_GBps_of_bytes_read": bytes_lvl0_one / (lvl0_one * 1e-3) / 1e9}
            del one_sets
            extras["train_step"] = train_step_leg(A, dev_sets, pose0, B, max(3, min(args.steps, 20)), dev)
            if args.workload == "tum":
                del streamer
                extras["vga480x640"] = vga_leg(A, rank, dev, args)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if peaks else "fallback 6650 GB/s"
    achieved = bytes_lvl0 / (lvl0_ms * 1e-3) / 1e9
    value = world * B * args.steps / (ms * 1e-3)
    out = {
        "metric": "frame-pair GN solves/sec", "value": value, "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "pairs_per_gpu_per_step": B, "feature_channels": C,
                   "resolution": f"{H}x{W}", "levels": N_LEVELS, "iters_per_level": ITERS, "variant": "U_IC",
                   "remove_tru_sigma": True, "pdl": not args.no_pdl, "streams": max(1, args.streams),
                   "sobel": "fused" if args.fused_sobel else "materialised once per level",
                   "lookups": "plain loads" if (args.no_staged or not args.fused_sobel) else "footprint staged in shared memory (cp.async ring)",
                   "launch": "single cooperative launch for all levels and iterations" if single else
                             ("one launch per iteration, each solve replayed from a CUDA graph" if use_graphs else "one launch per iteration"),
                   "l2": f"inputs rotate over {N_SETS} resident sets of {set_bytes / 1e6:.0f} MB each (> 126 MB L2)",
                   "algorithmic_bytes_per_step": bytes_step},
        "step_hbm_frac": bytes_step / (ms / args.steps * 1e-3) / 1e9 / peak,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": (328.3e6 if not args.fused_sobel else 184.0e6 if args.no_staged else 187.6e6)
                     if args.workload == "tum" else None,
                     "traffic_source": "ncu --set full dram__bytes_read+write per launch: profiles/r1h_uic_iter_staged_kernel_level0.txt "
                                       "(staged, default) / r1_uic_iter_kernel_level0.txt (--no-staged) / r1c_* (--materialised)",
                     "kernel": ("uic_iter_px_kernel<8,true>" if not args.fused_sobel else "uic_iter_kernel<8,true>" if args.no_staged
                                else "uic_iter_staged_kernel<true,false,160,120,false>") + " at the finest level",
                     "algorithmic_bytes_per_launch": bytes_lvl0, "launch_ms": lvl0_ms,
                     "all_launch_ms": [round(x, 4) for x in per_launch], "peak_source": peak_src,
                     "how": ("%globaltimer stamps at the iteration boundaries inside the single cooperative launch"
                             if single else "CUDA events around every launch") + " (dpft_uic_forward_timed), separate pass after the timed region"},
        "e2e": {"value": world * B * n_e2e / (ms_e2e * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": set_bytes,
                "d2h_bytes_per_step": B * 12 * 4, "steps": n_e2e, "ms_per_step": ms_e2e / n_e2e},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks.summary(),
        "extra": extras,
    }
    if not args.no_cpu_baseline:
        n_cpu = min(B, 16 if args.workload == "tum" else 1)
        v, threads, times = cpu_sample(data["levels"], data["R0"], data["t0"], n_cpu)
        out["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": threads, "kind": "port",
                               "sample": f"first {n_cpu} pairs of the step's batch, best of {len(times)} runs "
                                         f"({sum(times):.1f} s of CPU work)"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
