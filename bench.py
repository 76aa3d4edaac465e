#!/usr/bin/env python
"""Benchmark of the trust-region inverse-compositional solver path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload tum|vga|train|deepic|icp|tracker]

A *step* is one coarse-to-fine solve (4 pyramid levels x 3 Gauss-Newton iterations, U_IC with --remove_tru_sigma as
in every script the reference ships) of ONE batch of synthetic frame pairs -- 64 pairs at 120x160 (BASELINE config 2,
the default), 16 live frames at 480x640 against a resident keyframe (config 3, --workload vga), or one training step
of the patched reference tracker (config 4, --workload train).  --workload deepic (config 1: run_example.py's DeepIC
tracker, one pair per call) and --workload icp (config 5: --combine_ICP, batch 64) time the whole LeastSquareTracking
forward with the solver levels swapped by patch_tracker, next to the unpatched reference on the same GPU and on the
host cores.  The metric is frame-pair GN solves per second; rank 0 prints ONE JSON line.

  config     WHAT is measured (workload, sizes, the L2 rule): the same dict in both arms (the driver compares them)
  how        how this arm runs it (call sizes, streams, graph replay, which kernels, the uncertainty format)
  value      whole-job throughput with the inputs already in HBM, through the package's public BatchedSolver: calls
             of `batches_per_call` batches (each batch keeps its own batch-global sigma extremes, i.e. the results
             of separate reference calls) rotate over `streams` CUDA streams; K steps = K batches exactly.
             `latency_ms` next to it is ONE batch alone on one stream.  The device-timed region is bracketed by CUDA events
             on the main stream around exactly K steps; a 0.4 ms spin kernel is enqueued right before the start event, so
             the first call is already queued when the region starts (at 20 steps the region is one 4 ms call, and the
             host's submission latency would otherwise sit inside it; `how.timed_region`).
  e2e        same metric with HOST (pinned) buffers: every step uploads its batch and reads its poses back inside
             the timed region (HostStreamSolver, 3 device buffers).  The host holds what the reference's encoder
             emits: ONE uncertainty map per frame (alg:1425-1427 repeats it on the device).
  roofline   dominant kernel = the finest-level work-queue launch (uic_queue_kernel: 3 iterations of
             20 batches per launch, what a call of the timed region launches; in such calls the 60x80 and 30x40 levels run
             as work-queue launches of their own as well, algorithms.default_queue_levels); achieved = algorithmic bytes (4C+2)*4*H*W*B per iteration / CUDA-event time
  parity     this run's own results against the CPU oracle on the step's first batch (twist, J^T W J, mask flips)
  cpu_baseline  the reference itself (baseline/_ref, installed from /root/reference) on the host cores, full batches

N > 1 is launched by torchrun (one rank per GPU); frame pairs are independent, so every rank solves its own
batches and there is no collective on the data path (weak scaling).  --workload train adds the one collective of
this code base: the NCCL all-reduce of the encoder gradients (deep_prob_feature_track_b200/ddp.py).
"""
from __future__ import annotations

import argparse
import json
import os
import re
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
    # BASELINE.json configs[2]: 480x640 keyframe VO; a step = 16 live frames tracked against the resident keyframe
    "vga": dict(name="vga480x640_kf_b16_c8_uic_4lvl_x3it", B=16, C=8, H=480, W=640),
    # BASELINE.json configs[3]: training step (forward + backward + all-reduce + optimizer) at 120x160, batch 64 per GPU
    "train": dict(name="tum120x160_b64_c8_uic_trainstep", B=64, C=8, H=120, W=160),
    # BASELINE.json configs[0]: run_example.py's DeepIC tracker (IC solver, convolutional M-estimator, residual-volume
    # damping, run_example.py:84-92), one 120x160 RGB-D pair per call; the whole LeastSquareTracking forward
    "deepic": dict(name="tum120x160_b1_deepic_run_example", B=1, C=1, H=120, W=160,
                   flags=["--encoder_name", "ConvRGBD2", "--mestimator", "MultiScale2w", "--solver", "Direct-ResVol", "--uncertainty", "None"],
                   variant="IC (DeepIC: MultiScale2w M-estimator, Direct-ResVol)"),
    # BASELINE.json configs[4]: joint feature-metric + ICP refinement (scripts/train_tum_feature_icp.sh / eval_tum_feature_icp.sh),
    # batch 64; the whole LeastSquareTracking forward
    "icp": dict(name="tum120x160_b64_c8_uic_icp", B=64, C=8, H=120, W=160, flags="EVAL_TUM+combine_ICP",
                variant="U_IC + point-to-plane ICP term (--combine_ICP, constant scaler)"),
    # BASELINE.json configs[1] read literally ("feature pyramid + uncertainty-weighted trust-region IC solve"): the whole
    # LeastSquareTracking eval forward of scripts/eval_tum_rgbd.sh at batch 64, encoder included
    "tracker": dict(name="tum120x160_b64_c8_uic_tracker", B=64, C=8, H=120, W=160, flags="EVAL_TUM",
                    variant="U_IC, whole LeastSquareTracking forward (eval_tum_rgbd.sh)"),
}
TRACKER_WORKLOADS = ("deepic", "icp", "tracker")


def tracker_flags(wl):
    from baseline import reference as REF
    if wl["flags"] == "EVAL_TUM":
        return list(REF.EVAL_TUM_FLAGS)
    return REF.EVAL_TUM_FLAGS + ["--combine_ICP"] if wl["flags"] == "EVAL_TUM+combine_ICP" else list(wl["flags"])
N_LEVELS, ITERS = 4, 3
ROOFLINE_BATCHES = 20         # batches per launch of the roofline pass: what a call of the timed region launches (and what the
                              # committed ncu capture ran; at 8 per launch the tail of the launch weighs more: 45.9 against 44 us
                              # per batch-iteration, profiles/r2/r2d_uic_queue_kernel_onemap_level0_G8.txt)
NCU_SUMMARY = os.path.join(ROOT, "profiles", "r2", "r2f_uic_queue_kernel_onemap_level0_G20.txt")


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


def base_config(wl):
    """The part of `config` both arms print (the driver compares it)."""
    return {"workload": wl["name"], "pairs_per_gpu_per_step": wl["B"], "feature_channels": wl["C"],
            "resolution": f"{wl['H']}x{wl['W']}", "levels": N_LEVELS, "iters_per_level": ITERS, "variant": wl.get("variant", "U_IC"),
            "remove_tru_sigma": "flags" not in wl or str(wl["flags"]).startswith("EVAL_TUM")}


def l2_note(wl, workload, args):
    """How the timed region keeps a step from finding its inputs in the 126 MB L2 (same text in both arms: it is part
    of `config`, which the driver compares)."""
    B, C, H, W = wl["B"], wl["C"], wl["H"], wl["W"]
    S = max(1, args.streams)
    if workload == "tum":
        n_stack = max(max(1, args.batches_per_call), ROOFLINE_BATCHES)
        per_batch = sum(((4 * C + 2) * (H >> l) * (W >> l) + 4) * 4 * B for l in range(N_LEVELS))
        return (f"every stream owns a set of {n_stack} batches ({per_batch / 1e6:.0f} MB each, > 126 MB L2 per call); "
                f"sets differ in addresses and pair order")
    if workload == "vga":
        per_set = sum(((2 * C + 1) * (H >> l) * (W >> l) + 4) * 4 * B for l in range(N_LEVELS))
        return f"live frames rotate over {max(2, S)} resident sets of {per_set / 1e6:.0f} MB (> 126 MB L2)"
    per_batch = B * 8 * H * W * 4 + 16 * B
    if workload == "train":
        return f"2 RGB-D batches of {per_batch / 1e6:.0f} MB; every step's features, gradients and workspaces are new tensors"
    n_sets = min(max(4, -(-140_000_000 // (B * 8 * H * W * 4))), 64)
    return f"RGB-D batches rotate over {n_sets} resident sets ({n_sets * per_batch / 1e6:.0f} MB); every step's features are new tensors"


def workload_config(wl, workload, args):
    """`config` of the JSON line: what is measured -- identical in both arms.  How this arm runs it goes to `how`."""
    return dict(base_config(wl), l2=l2_note(wl, workload, args))


# ----------------------------------------------------------------------------------------------- CPU reference
def cpu_reference_runner(wl, workload, seed):
    """(step(), pairs_per_step, kind, description): the reference's own implementation of the path on the host cores.
    baseline/_ref (the installed reference) when it is there, else the oracle port."""
    from deep_prob_feature_track_b200.synthetic import make_frame_pairs
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    B, C, H, W = wl["B"], wl["C"], wl["H"], wl["W"]
    try:
        from baseline import reference as REF
        have_ref = REF.available()
    except Exception:
        have_ref = False
    if workload == "train":
        n = 8
        if not have_ref:
            raise SystemExit("--workload train --impl reference needs baseline/_ref (python baseline/install_reference.py)")
        REF.modules()                         # puts baseline/_ref/code on the import path
        import models.criterions as crit
        net = REF.make_tracker(REF.TRAIN_TUM_FLAGS).train()
        img0, img1, d0, d1, K = REF.synthetic_rgbd(n, H, W, seed=seed)
        R_gt = torch.eye(3).repeat(n, 1, 1)
        t_gt = torch.tensor([[0.01, -0.005, 0.002]]).repeat(n, 1)
        opt = torch.optim.Adam(net.parameters(), lr=1e-5, weight_decay=4e-4)

        def step():
            opt.zero_grad()
            Rs, ts = net(img0, img1, d0, d1, K)
            loss = crit.compute_RT_EPE_loss(Rs, ts, R_gt, t_gt, d0, K, invalid=(d0 < 0.1)).mean() * 1e2
            loss.backward()
            torch.nn.utils.clip_grad_norm_(net.parameters(), 5.0)
            opt.step()
        return step, n, "reference", (f"training step of the reference's LeastSquareTracking (train.py:117-192) on {n} of the "
                                      f"{B} pairs of a step")
    if workload in TRACKER_WORKLOADS:
        if not have_ref:
            raise SystemExit(f"--workload {workload} needs the reference's networks: baseline/_ref is missing (python baseline/install_reference.py)")
        n = min(B, 16)
        net = REF.make_tracker(tracker_flags(wl)).eval()
        rgbd = REF.synthetic_rgbd(n, H, W, seed=seed)

        def step():
            with torch.no_grad():
                net(*rgbd)
        return step, n, "reference", (f"forward of the reference's LeastSquareTracking ({' '.join(tracker_flags(wl))}) on {n} of the {B} pairs of a step")
    n = B if workload == "tum" else 1
    data = make_frame_pairs(n, C, H, W, seed=seed, n_levels=N_LEVELS)
    pose = [data["R0"], data["t0"].view(n, 3, 1)]
    if have_ref:
        net = REF.make_tracker().eval()

        def step():
            with torch.no_grad():
                REF.solver_chain(net, data["levels"], pose)
        return step, n, "reference", (f"tr_update3..0 of the reference's LeastSquareTracking (LeastSquareTracking.py:345-446) on "
                                      f"{n} of the {B} pairs of a step")
    from oracle import ic_oracle as O

    def step():
        with torch.no_grad():
            O.track_pyramid(data["levels"], (data["R0"], data["t0"]), iters=ITERS, remove_tru_sigma=True,
                            sampler="grid_sample", reduction="bmm")
    return step, n, "port", f"oracle port of the reference op chain on {n} of the {B} pairs of a step (baseline/_ref missing)"


def cpu_sample(wl, workload, seed, min_seconds=10.0):
    """Mean throughput of full steps of the CPU reference over >= min_seconds (the statistic of the reference arm)."""
    step, n, kind, what = cpu_reference_runner(wl, workload, seed)
    step()
    times = []
    t_begin = time.perf_counter()
    while len(times) < 2 or time.perf_counter() - t_begin < min_seconds:
        t = time.perf_counter()
        step()
        times.append(time.perf_counter() - t)
    return {"value": n * len(times) / sum(times), "unit": "pairs/s", "cores": os.cpu_count() or 1, "kind": kind,
            "sample": f"{what}; mean of {len(times)} steps after 1 warm-up ({sum(times):.1f} s of CPU work)"}


def reference_arm(args, wl):
    step, n, kind, what = cpu_reference_runner(wl, args.workload, 1234)
    for _ in range(args.warmup):
        step()
    t = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t
    v = n * args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": "frame-pair GN solves/sec", "value": v, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(wl, args.workload, args),
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": os.cpu_count() or 1, "kind": kind,
                         "sample": f"{what}; mean of {args.steps} steps after {args.warmup} warm-up"},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------------------------- helpers (GPU)
def event_time_ms(fn, main):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    fn()
    e1.record(main)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def gate(main, dev, us=400.0):
    """A spin kernel on `main` right before the start event of a device-timed region: while it runs the host enqueues the
    event and the first call, so the region starts with work already queued instead of with the host's submission latency
    (at the driver's 20 steps the tum region is ONE 4 ms call; 8 ranks on 16 vCPUs submitted it 0.1-0.3 ms late).  The
    events still bracket exactly the K steps; nothing of ours runs in the gate."""
    try:
        khz = getattr(torch.cuda.get_device_properties(dev), "clock_rate", 1965000)
        with torch.cuda.stream(main):
            torch.cuda._sleep(int(us * 1e-6 * khz * 1e3))
        return True
    except Exception:
        return False


def stacked_sets(gen_dev, n_sets, n_batches):
    """n_sets distinct allocations of n_batches batches each, composed from the generated batches rolled along the
    batch axis (distinct addresses and pair orders: no step finds another step's inputs in L2)."""
    sets = []
    for s in range(n_sets):
        parts = [(gen_dev[(s + j) % len(gen_dev)], (5 * j + 3 * s) % 64) for j in range(n_batches)]
        levels = [{k: torch.cat([torch.roll(p["levels"][l][k], sh, 0) if (sh and k != "K") else p["levels"][l][k]
                                 for p, sh in parts]).contiguous() for k in parts[0][0]["levels"][l]}
                  for l in range(N_LEVELS)]
        pose = (torch.cat([p["R0"] for p, _ in parts]), torch.cat([p["t0"] for p, _ in parts]))
        sets.append((levels, pose))
    return sets


def take(levels, pose, n):
    return [{k: v[:n] for k, v in lv.items()} for lv in levels], (pose[0][:n], pose[1][:n])


def ncu_traffic(path):
    """dram bytes read + written of the committed ncu --set full capture of the roofline kernel (per launch)."""
    try:
        txt = open(path).read()
        rd = re.search(r"dram__bytes_read\.sum\s+(\S+)\s+([0-9.]+)", txt)
        wr = re.search(r"dram__bytes_write\.sum\s+(\S+)\s+([0-9.]+)", txt)
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        return float(rd.group(2)) * scale[rd.group(1)] + float(wr.group(2)) * scale[wr.group(1)]
    except Exception:
        return None


def parity_block(A, batch_cpu, res_rows, lpi_occ):
    """This run's results on one batch against the CPU oracle: twist, J^T W J of every iteration, mask flips."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import frob_rel, twist_rel_err
    from oracle import ic_oracle as O
    trace = []
    with torch.no_grad():
        pose, _ = O.track_pyramid(batch_cpu["levels"], (batch_cpu["R0"], batch_cpu["t0"]), iters=ITERS, remove_tru_sigma=True,
                                  trace=trace, reduction="einsum")
    R, t = A.unpack_pose(res_rows["pose_hist"][-1].cpu())
    relA, relb, flips, pixels = 0.0, 0.0, 0, 0
    for i, tr in enumerate(trace):
        for it, rec in enumerate(tr):
            k = i * ITERS + it
            Ac, bc = A.unpack_system(res_rows["sys_hist"][k].cpu())
            relA, relb = max(relA, frob_rel(Ac, rec["A"])), max(relb, frob_rel(bc, rec["b"]))
            occ = lpi_occ[i][it].cpu()
            flips += int((occ != rec["occ"][:, 0].to(torch.uint8)).sum())
            pixels += occ.numel()
    n4 = 4
    return {"against": "oracle/ic_oracle.py (CPU restatement of the reference, pinned to reference-generated fixtures)",
            "pairs": int(R.shape[0]), "twist_rel_err_max": twist_rel_err(R, t, pose[0], pose[1]),
            "twist_rel_err_max_first4": twist_rel_err(R[:n4], t[:n4], pose[0][:n4], pose[1][:n4]),
            "pose_abs_err_max": max((R - pose[0]).abs().max().item(), (t - pose[1]).abs().max().item()),
            "JtWJ_frob_rel_max": relA, "JtWr_frob_rel_max": relb, "mask_flips": flips, "mask_pixels": pixels,
            "note": "masks of the first iteration are bit-exact (tests); later iterations start from poses that differ in "
                    "the last bits, so threshold-adjacent pixels may flip"}


# ----------------------------------------------------------------------------------------------- workloads (GPU)
def run_tum(args, wl, rank, world, dev, barrier, max_over_ranks):
    from deep_prob_feature_track_b200 import algorithms as A
    from deep_prob_feature_track_b200.batched import BatchedSolver, HostStreamSolver, pack_levels
    from deep_prob_feature_track_b200.synthetic import make_frame_pairs
    B, C, H, W = wl["B"], wl["C"], wl["H"], wl["W"]
    K, S = args.steps, max(1, args.streams)
    G = max(1, min(args.batches_per_call, K))
    bytes_step, bytes_lvl0 = algorithmic_bytes(B, C, H, W)
    main = torch.cuda.current_stream(dev)

    gen = [make_frame_pairs(B, C, H, W, seed=1234 + 100 * rank + i, n_levels=N_LEVELS) for i in range(4)]
    gen_dev = [{"levels": [{k: v.to(dev) for k, v in lv.items()} for lv in g["levels"]], "R0": g["R0"].to(dev),
                "t0": g["t0"].to(dev)} for g in gen]
    n_stack = max(G, ROOFLINE_BATCHES)
    sets = stacked_sets(gen_dev, S, n_stack)               # one stacked set per stream
    set_bytes = sum(v.numel() * 4 for lv in sets[0][0] for v in lv.values()) // n_stack
    solver = BatchedSolver(B, iters=ITERS, remove_tru_sigma=True, streams=S, device=dev, graphs=not args.no_graphs,
                           **({"queue_ctas": args.queue_ctas} if args.queue_ctas > 0 else {}))

    calls = [G] * (K // G) + ([K % G] if K % G else [])     # K steps exactly

    def run(call_sizes):
        out = None
        for i, g in enumerate(call_sizes):
            out = solver.submit(*take(*sets[i % S], B * g))
        return out

    # ---- value
    # warm-up: every call size of the timed region on every input set (function attributes, allocator, and the graph
    # of that set and size is captured here, not in the timed region), then whole calls up to at least W steps
    res = None
    for g in sorted(set(calls + [1]), reverse=True):
        for si in range(S):
            res = solver.submit(*take(*sets[si], B * g))
    done = S * sum(set(calls))
    if done < args.warmup:
        run([G] * ((args.warmup - done + G - 1) // G))
    solver.synchronize()
    res.raise_if_bad()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gated = gate(main, dev)
    e0.record(main)
    solver.wait_for(e0)
    res = run(calls)
    solver.join(main)
    e1.record(main)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1), dev)
    res.raise_if_bad()
    # kernels of one call: init of the replication flag and the sigma0 extremes, the replication check (which collects those
    # extremes on its way), init of poses and counters, 3 iterations of the 15x20 level; then
    # per level either 3 x 2 twin launches (resident 30x40 / staged 60x80 launch-per-iteration kernels) or, for the levels
    # that join the queue when the call is large enough (algorithms.default_queue_levels), queue init + its two twins
    def launches_per_call(g):
        ql = A.default_queue_levels([(H >> l, W >> l) for l in (3, 2, 1, 0)], g * B)
        return 1 + 2 + ITERS + (3 if ql >= 3 else 2 * ITERS) + (3 if ql >= 2 else 2 * ITERS) + 3

    # ---- latency of ONE batch alone (launch-per-iteration kernels, one stream)
    one = [take(*s, B) for s in sets]
    lat = []
    for i in range(3 + 10):
        t = event_time_ms(lambda: A.uic_solve(*one[i % S], iters=ITERS, remove_tru_sigma=True), main)
        if i >= 3:
            lat.append(t)
    latency_ms = statistics.median(lat)

    # ---- roofline: the finest-level work-queue launch, ROOFLINE_BATCHES batches per launch, CUDA events around it
    full = A.uic_solve(*take(*sets[0], B * ROOFLINE_BATCHES), iters=ITERS, remove_tru_sigma=True, group=B, queue=True)
    pose_l0 = A.unpack_pose(full.pose_hist[(N_LEVELS - 1) * ITERS])
    fine = [take(*s, B * ROOFLINE_BATCHES)[0][-1] for s in sets]

    def fine_launch(i, **kw):
        return A.uic_solve([fine[i % S]], pose_l0, iters=ITERS, remove_tru_sigma=True, group=B, queue=True, **kw)
    for i in range(3):
        fine_launch(i)
    torch.cuda.synchronize()
    call_ms = statistics.mean(event_time_ms(lambda i=i: fine_launch(i), main) for i in range(max(3, min(K, 10))))
    # the kernel alone: CUDA events recorded by the library right before / after the work-queue kernel's launch (with the
    # replication check on, around its two twin launches, of which the one that does not apply returns at once)
    timed = [fine_launch(i, timed=True) for i in range(max(3, min(K, 10)))]
    ev = [r.queue_kernel_ms[0] for r in timed]
    launch_ms = statistics.mean(ev)
    stamps = timed[0].launch_ms
    assert (fine_launch(0).pose_hist[-1] - full.pose_hist[-1]).abs().max().item() < 1e-5   # the timed launch does the real work
    # the same level with the check off: the C-map tile routine (what sigma tensors with independent channels run)
    cmap = [fine_launch(i, timed=True, tuning=dict(sigma_detect=1)) for i in range(max(3, min(K, 10)))]
    cmap_ms = statistics.mean(r.queue_kernel_ms[0] for r in cmap)
    assert (cmap[0].pose_hist[-1] - full.pose_hist[-1]).abs().max().item() < 1e-5

    # ---- parity of this very run (first batch of the first set) against the CPU oracle
    parity = None
    if rank == 0 and not args.no_parity:
        first = A.uic_solve(*take(*sets[0], B * G), iters=ITERS, remove_tru_sigma=True, group=B, **solver.solve_kw)
        lpi = A.uic_solve(*one[0], iters=ITERS, remove_tru_sigma=True, queue=False, want_occ=True)
        torch.cuda.synchronize()
        parity = parity_block(A, gen[0], {"pose_hist": first.pose_hist[:, :B], "sys_hist": first.sys_hist[:, :B]}, lpi.occ)

    # ---- e2e: every step's batch comes from pinned host memory and its poses go back
    host_levels = [dict(lv, s0=lv["s0"][:, :1].contiguous(), s1=lv["s1"][:, :1].contiguous()) for lv in gen[0]["levels"]]
    assert all(torch.equal(lv["s0"], h["s0"].expand_as(lv["s0"])) for lv, h in zip(gen[0]["levels"], host_levels))
    host_flat, layout = pack_levels(host_levels, pin=True)
    host_flats = [host_flat] + [host_flat.clone().pin_memory() for _ in range(2)]
    pose0 = (gen_dev[0]["R0"], gen_dev[0]["t0"])
    streamer = HostStreamSolver(layout, host_flat.numel(), N_LEVELS, B, dev,
                                lambda lv: A.uic_solve(lv, pose0, iters=ITERS, remove_tru_sigma=True), depth=3)
    n_e2e = K
    streamer.run([host_flats[i % 3] for i in range(3)])
    barrier()
    e0.record(main)
    out_host = streamer.run([host_flats[i % 3] for i in range(n_e2e)])
    e1.record(main)
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1), dev)
    torch.cuda.synchronize()
    e2e_ok = (out_host.to(dev) - A.uic_solve(gen_dev[0]["levels"], pose0, iters=ITERS, remove_tru_sigma=True).pose_hist[-1]).abs().max().item()

    extras = {}
    if not args.no_extras and rank == 0:
        # the same solve with the uncertainty passed as the ONE map per frame the reference's encoder emits
        one_sigma = [([dict(lv, s0=lv["s0"][:, :1].contiguous(), s1=lv["s1"][:, :1].contiguous()) for lv in s[0]], s[1]) for s in sets[:1]]
        tmp = BatchedSolver(B, iters=ITERS, remove_tru_sigma=True, streams=1, device=dev)
        for _ in range(2):
            tmp.submit(*take(*one_sigma[0], B * G))
        tmp.synchronize()
        e0.record(main)
        tmp.wait_for(e0)
        n_sb = max(2, min(4, K // G))
        for _ in range(n_sb):
            tmp.submit(*take(*one_sigma[0], B * G))
        tmp.join(main)
        e1.record(main)
        torch.cuda.synchronize()
        extras["single_sigma_map"] = {"what": "sigma0 / sigma1 given as (B,1,h,w) instead of repeated to C channels "
                                              "(DPFT_SIGMA_BROADCAST); same results, one stream",
                                      "pairs_per_s": B * G * n_sb / (e0.elapsed_time(e1) * 1e-3)}
        del one_sigma, tmp
        extras["train_step_solver_only"] = train_solver_leg(A, one, B, dev, main)

    peak, peak_src = measured_peak()
    achieved = ITERS * bytes_lvl0 * ROOFLINE_BATCHES / (launch_ms * 1e-3) / 1e9
    # what the one-map routine reads per pixel: x0, x1 (C channels), ONE sigma0 / sigma1 map, both inverse depths
    bytes_read_lvl0 = (2 * C + 4) * 4 * H * W * B
    achieved_read = ITERS * bytes_read_lvl0 * ROOFLINE_BATCHES / (launch_ms * 1e-3) / 1e9
    achieved_cmap = ITERS * bytes_lvl0 * ROOFLINE_BATCHES / (cmap_ms * 1e-3) / 1e9
    traffic = ncu_traffic(NCU_SUMMARY)
    return {
        "value": world * B * K / (ms * 1e-3), "ms_per_step": ms / K,
        "config": dict(base_config(wl), batches_per_call=G, streams=S, call_sizes=calls,
                       api="deep_prob_feature_track_b200.batched.BatchedSolver.submit",
                       timed_region=("CUDA events on the main stream around exactly K steps; a 0.4 ms spin kernel precedes the start event so "
                                     "that the first call is already queued when the region starts (no host submission latency inside)"
                                     if gated else "CUDA events on the main stream around exactly K steps"),
                       cuda_graphs=f"{solver.replays} of {solver.calls} calls replayed from a captured graph (BatchedSolver(graphs=True): "
                                   "one graph per input set, captured during warm-up)",
                       coarse_levels="one launch per Gauss-Newton iteration (uic_iter_staged_kernel / uic_iter_kernel), all batches of a call in one grid",
                       finest_level="one work-queue launch for its 3 iterations (uic_queue_kernel, per-pair dependencies); in calls of "
                                    "4 / 6 batches or more the 60x80 / 30x40 levels run as work-queue launches of their own as well "
                                    "(algorithms.default_queue_levels; the 30x40 level on the staged routine's narrow form)",
                       sigma="(B,C,H,W) tensors holding C copies of one map per frame, as the reference's encoder emits them; a device-side "
                             "check per call finds that (one read of the sigma tensors, inside the timed region) and the one-map tile "
                             "routines run (options.sigma_detect)",
                       sigma_extremes="per batch of 64 (options.group): the results of separate reference calls",
                       l2=l2_note(wl, "tum", args),
                       algorithmic_bytes_per_step=bytes_step),
        "latency_ms": latency_ms,
        "latency_note": "one batch of 64 pairs alone on one stream (launch-per-iteration kernels): ms per solve",
        "step_hbm_frac": bytes_step / (ms / K * 1e-3) / 1e9 / peak,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic,
                     "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of the same launch "
                                       f"({ROOFLINE_BATCHES} batches): {os.path.relpath(NCU_SUMMARY, ROOT)}",
                     "kernel": "uic_queue_kernel<true,true,false,0,0,1>: the finest level (120x160) of "
                               f"{ROOFLINE_BATCHES} batches of {B} pairs, its {ITERS} Gauss-Newton iterations in one launch, ONE-MAP tile "
                               "routine: the (B,C,H,W) sigma tensors of the workload are C copies of one map (as the reference's encoder "
                               "emits them, SURVEY 8d), sigma_replication_kernel finds that on the device and this twin of the launch runs",
                     "algorithmic_bytes_per_launch": ITERS * bytes_lvl0 * ROOFLINE_BATCHES, "launch_ms": launch_ms,
                     "algorithmic_note": "achieved / frac use SURVEY 8(d)'s (4C+2)*4 bytes per pixel and iteration; this kernel reads "
                                         "(2C+4)*4 of them (x0, x1, one sigma0 / sigma1 map, inverse depths): see bytes_read_*",
                     "bytes_read_per_launch": ITERS * bytes_read_lvl0 * ROOFLINE_BATCHES,
                     "bytes_read_GBps": achieved_read, "bytes_read_frac": achieved_read / peak,
                     "c_map_kernel": {"kernel": "uic_queue_kernel<true,false,false,160,120,1> (options.sigma_detect = 1: every sigma "
                                                "channel read, what tensors with independent channels run)",
                                      "launch_ms": cmap_ms, "achieved": achieved_cmap, "frac": achieved_cmap / peak,
                                      "ncu": "profiles/r2/r2b_uic_queue_kernel_level0_G20.txt"},
                     "launch_ms_all": [round(x, 4) for x in ev],
                     "iteration_ms_device_stamps": [round(x, 4) for x in stamps],
                     "per_batch_iteration_us": launch_ms * 1e3 / (ITERS * ROOFLINE_BATCHES),
                     "peak_source": peak_src,
                     "call_ms_with_helper_launches": call_ms,
                     "how": "CUDA events recorded on the launching stream right before and right after the kernel's launch "
                            "(dpft_uic_options.queue_kernel_ms), separate pass after the timed region; call_ms adds the "
                            "launches around it (sigma0 extremes of the level, queue init); iteration_ms_device_stamps are "
                            "%globaltimer stamps of the iteration completions inside the kernel"},
        "parity": parity,
        "e2e": {"value": world * B * n_e2e / (ms_e2e * 1e-3), "unit": "pairs/s",
                "h2d_bytes_per_step": streamer.h2d_bytes_per_step, "d2h_bytes_per_step": streamer.d2h_bytes_per_step,
                "steps": n_e2e, "ms_per_step": ms_e2e / n_e2e,
                "what": "pinned host batch (x0, x1 with 8 channels, ONE sigma map per frame as the reference's encoder emits it, "
                        "inverse depths, K) -> 3 device buffers -> solve -> poses in pinned memory, every step",
                "pose_check_max_abs": e2e_ok},
        "gpu_launches": sum(launches_per_call(g) for g in calls),
        "extra": extras,
    }


def train_solver_leg(A, one, B, dev, main):
    """Solver part of a training step alone (forward + backward through all 12 iterations, linear loss)."""
    leaves = [[{k: (v.clone().requires_grad_(True) if k in ("x0", "x1", "s0", "s1") else v) for k, v in lv.items()}
               for lv in s[0]] for s in one[:2]]
    R = one[0][1][0].clone().requires_grad_(True)
    t = one[0][1][1].clone().requires_grad_(True)

    def step(i):
        for lv in leaves[i % len(leaves)]:
            for k in ("x0", "x1", "s0", "s1"):
                lv[k].grad = None
        outs = A.uic_track(leaves[i % len(leaves)], (R, t), iters=ITERS, remove_tru_sigma=True, check=False)
        sum(Rl.sum() + tl.sum() for Rl, tl, _ in outs).backward()
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    n = 10
    ms = event_time_ms(lambda: [step(i) for i in range(n)], main) / n

    # the backward of the finest level alone (its 3 iterations): what uic_bwd_px_kernel and the kernels around it cost
    C, H, W = (int(v) for v in leaves[0][-1]["x0"].shape[1:])
    bwd = []
    for i in range(3 + 5):
        lv = leaves[i % len(leaves)][-1]
        for k in ("x0", "x1", "s0", "s1"):
            lv[k].grad = None
        outs = A.uic_track([lv], (R, t), iters=ITERS, remove_tru_sigma=True, check=False)
        loss = sum(Rl.sum() + tl.sum() for Rl, tl, _ in outs)
        t_b = event_time_ms(loss.backward, main)
        if i >= 3:
            bwd.append(t_b)
    bwd_ms = statistics.median(bwd)
    bwd_bytes = ITERS * (12 * C + 2) * 4 * H * W * B
    peak, _ = measured_peak()
    return {"what": "solver forward+backward alone (autograd through 4 levels x 3 iterations), one GPU", "ms_per_step": ms,
            "pairs_per_s": B / (ms * 1e-3),
            "backward_finest_level": {
                "what": "dpft_uic_backward of the 120x160 level alone (3 iterations: unit Sobel maps, 3 x uic_bwd_px_kernel, 3 x pose_bwd, "
                        "the Sobel adjoint, zero fills), CUDA events around autograd's backward",
                "ms": bwd_ms, "algorithmic_bytes": bwd_bytes,
                "algorithmic_note": "(12C+2)*4 bytes per pixel and iteration: x0, sigma0, their unit Sobel maps, x1 / sigma1 lookups and "
                                    "the gradient maps written once; the kernel moves 1.9x that (each iteration reads and rewrites the six "
                                    "accumulated maps per channel, profiles/r2/r2b_bwd_variants.txt)",
                "achieved_GBps": bwd_bytes / (bwd_ms * 1e-3) / 1e9, "frac": bwd_bytes / (bwd_ms * 1e-3) / 1e9 / peak}}


def run_vga(args, wl, rank, world, dev, barrier, max_over_ranks):
    """BASELINE config 3: keyframe VO at 480x640.  A sequence's keyframe stays resident; a step tracks 16 live frames
    against it (per-frame sigma extremes = the semantics of kf_vo.py's B = 1 calls)."""
    from deep_prob_feature_track_b200 import algorithms as A
    from deep_prob_feature_track_b200.batched import HostStreamSolver, pack_levels
    from deep_prob_feature_track_b200.synthetic import make_frame_pairs
    B, C, H, W = wl["B"], wl["C"], wl["H"], wl["W"]
    K, S = args.steps, max(1, args.streams)
    bytes_step, bytes_lvl0 = algorithmic_bytes(B, C, H, W)
    main = torch.cuda.current_stream(dev)
    data = make_frame_pairs(B, C, H, W, seed=99 + rank, n_levels=N_LEVELS, motion=0.05)
    key = [{k: lv[k][:1].to(dev).contiguous() for k in ("x0", "s0", "invD0")} for lv in data["levels"]]
    live_cpu = [{k: lv[k] for k in ("x1", "s1", "invD1", "K")} for lv in data["levels"]]
    lives = [[{k: (torch.roll(v, s, 0) if k != "K" else v).to(dev).contiguous() for k, v in lv.items()} for lv in live_cpu]
             for s in range(max(2, S))]
    pose0 = (data["R0"].to(dev), data["t0"].to(dev))
    tracker = A.KeyframeTracker(key, iters=ITERS, remove_tru_sigma=True)
    streams = [torch.cuda.Stream(device=dev) for _ in range(S)]

    def run(n):
        out = None
        for i in range(n):
            with torch.cuda.stream(streams[i % S]):
                out = tracker.track(lives[i % len(lives)], pose0)
        return out
    res = run(max(args.warmup, S))
    torch.cuda.synchronize()
    res.raise_if_bad()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gate(main, dev)
    e0.record(main)
    for st in streams:
        st.wait_event(e0)
    res = run(K)
    for st in streams:
        evd = torch.cuda.Event()
        evd.record(st)
        main.wait_event(evd)
    e1.record(main)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1), dev)
    res.raise_if_bad()

    # latency of ONE live frame (kf_vo.py's callback: B = 1 per call)
    live1 = [{k: v[:1] for k, v in lv.items()} for lv in lives[0]]
    p1 = (pose0[0][:1], pose0[1][:1])
    for _ in range(3):
        tracker.track(live1, p1)
    torch.cuda.synchronize()
    lat1 = statistics.median(event_time_ms(lambda: tracker.track(live1, p1), main) for _ in range(10))

    # roofline: the finest-level work-queue launch of one step
    full = tracker.track(lives[0], pose0)
    pose_l0 = A.unpack_pose(full.pose_hist[(N_LEVELS - 1) * ITERS])
    kfine = A.KeyframeTracker(key[-1:], iters=ITERS, remove_tru_sigma=True)
    for i in range(3):
        kfine.track([lives[i % len(lives)][-1]], pose_l0)
    torch.cuda.synchronize()
    call_ms = statistics.mean(event_time_ms(lambda i=i: kfine.track([lives[i % len(lives)][-1]], pose_l0), main)
                              for i in range(max(3, min(K, 10))))
    # the kernel alone, as in the tum workload: CUDA events recorded by the library right before / after the work-queue
    # kernel's launch (around its two twin launches, of which the one that does not apply returns at once)
    timed = [kfine.track([lives[i % len(lives)][-1]], pose_l0, timed=True) for i in range(max(3, min(K, 10)))]
    launch_ms = statistics.mean(r.queue_kernel_ms[0] for r in timed)
    assert (timed[0].pose_hist[-1] - full.pose_hist[-1]).abs().max().item() < 1e-5   # the timed launch does the real work

    # parity against the oracle, frame by frame (B = 1 semantics), first 2 live frames
    parity = None
    if rank == 0 and not args.no_parity:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from helpers import twist_rel_err
        from oracle import ic_oracle as O
        errs = []
        Rc, tc = (x.cpu() for x in full.pose)
        for b in range(2):
            lv1 = [dict(x0=lv["x0"][:1], s0=lv["s0"][:1], invD0=lv["invD0"][:1], **{k: v[b:b + 1] for k, v in lc.items()})
                   for lv, lc in zip(data["levels"], live_cpu)]
            with torch.no_grad():
                (R, t), _ = O.track_pyramid(lv1, (data["R0"][b:b + 1], data["t0"][b:b + 1]), iters=ITERS, remove_tru_sigma=True)
            errs.append((twist_rel_err(Rc[b:b + 1], tc[b:b + 1], R, t), max((Rc[b] - R[0]).abs().max().item(), (tc[b] - t[0]).abs().max().item())))
        parity = {"against": "oracle/ic_oracle.py, one B = 1 call per live frame (kf_vo.py:156-166)", "pairs": 2,
                  "twist_rel_err_max": max(e[0] for e in errs), "pose_abs_err_max": max(e[1] for e in errs)}

    # e2e: the live frames of every step come from pinned host memory (the keyframe side stays resident)
    host_live = [dict(lv, s1=lv["s1"][:, :1].contiguous()) for lv in live_cpu]
    key_sb = [dict(kf, s0=kf["s0"][:, :1].contiguous()) for kf in key]
    tracker_sb = A.KeyframeTracker(key_sb, iters=ITERS, remove_tru_sigma=True)
    host_flat, layout = pack_levels(host_live, pin=True)
    host_flats = [host_flat, host_flat.clone().pin_memory()]
    streamer = HostStreamSolver(layout, host_flat.numel(), N_LEVELS, B, dev, lambda lv: tracker_sb.track(lv, pose0), depth=3)
    streamer.run(host_flats)
    barrier()
    e0.record(main)
    streamer.run([host_flats[i % 2] for i in range(K)])
    e1.record(main)
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1), dev)

    peak, peak_src = measured_peak()
    achieved = ITERS * bytes_lvl0 / (launch_ms * 1e-3) / 1e9
    return {
        "value": world * B * K / (ms * 1e-3), "ms_per_step": ms / K,
        "config": dict(base_config(wl), streams=S, keyframe="one resident keyframe per sequence (x0, sigma0, invD0 with batch size 1), "
                       f"{B} live frames per step, sigma extremes per frame (DPFT_SHARED_KEYFRAME | DPFT_PAIRWISE_EXTREMES)",
                       api="deep_prob_feature_track_b200.algorithms.KeyframeTracker.track",
                       l2=l2_note(wl, "vga", args),
                       algorithmic_bytes_per_step=bytes_step),
        "latency_ms": lat1, "latency_note": "ONE live frame against the keyframe (B = 1, kf_vo.py's per-frame call): ms per solve",
        "step_hbm_frac": bytes_step / (ms / K * 1e-3) / 1e9 / peak,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(os.path.join(ROOT, "profiles", "r2", "r2d_uic_queue_kernel_vga_level0.txt")) if B == 16 else None,
                     "traffic_source": "ncu --set full of the same launch at 16 frames: profiles/r2/r2d_uic_queue_kernel_vga_level0.txt",
                     "kernel": f"uic_queue_kernel<true,true,false,0,0,1>: the finest level (480x640) of {B} live frames, {ITERS} iterations in one launch "
                               "(the one-map twin: the sigma tensors are C copies of one map, found on the device)",
                     "algorithmic_bytes_per_launch": ITERS * bytes_lvl0, "launch_ms": launch_ms, "peak_source": peak_src,
                     "note": "algorithmic bytes count the keyframe side once per pair (SURVEY 8d); it is shared by the 16 frames, so DRAM reads less",
                     "call_ms_with_helper_launches": call_ms,
                     "how": "CUDA events recorded on the launching stream right before and right after the kernel's launch "
                            "(dpft_uic_options.queue_kernel_ms), separate pass after the timed region; call_ms adds the replication "
                            "check, the sigma0 extremes, the queue init and the idle twin"},
        "parity": parity,
        "e2e": {"value": world * B * K / (ms_e2e * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": streamer.h2d_bytes_per_step,
                "d2h_bytes_per_step": streamer.d2h_bytes_per_step, "steps": K, "ms_per_step": ms_e2e / K,
                "what": "pinned host live frames (x1 with 8 channels, one sigma map, inverse depth, K) -> device -> track against the "
                        "resident keyframe -> poses in pinned memory, every step"},
        "gpu_launches": (2 + 3 * ITERS + 4) * K,
        "extra": {},
    }


def run_tracker(args, wl, rank, world, dev, barrier, max_over_ranks):
    """BASELINE configs 1 and 5: the whole LeastSquareTracking forward (the reference's encoder and in-loop networks on
    cuDNN, its tr_update0..3 swapped by patch_tracker) on device-resident RGB-D batches, next to the unpatched
    reference on the same GPU and on the host cores."""
    import copy
    from deep_prob_feature_track_b200 import algorithms as A
    from baseline import reference as REF
    if not REF.available():
        raise SystemExit(f"--workload {args.workload} needs the reference's networks: baseline/_ref is missing (python baseline/install_reference.py)")
    B, H, W = wl["B"], wl["H"], wl["W"]
    K = args.steps
    main = torch.cuda.current_stream(dev)
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = REF.make_tracker(tracker_flags(wl), seed=0).to(dev).eval()
    # (U_IC without the ICP term: all four levels in one solver call; the other trackers level by level)
    ours = A.patch_tracker(copy.deepcopy(ref), fused_forward=(args.workload == "tracker" and not args.level_by_level)).eval()
    n_sets = max(4, -(-140_000_000 // (B * 8 * H * W * 4)))      # inputs of the rotation exceed the 126 MB L2 ... where B allows
    n_sets = min(n_sets, 64)
    host = [tuple(x.pin_memory() for x in REF.synthetic_rgbd(B, H, W, seed=70 + 100 * rank + i)) for i in range(n_sets)]
    batches = [tuple(x.to(dev) for x in hb) for hb in host]

    def run(net, n):
        out = None
        with torch.no_grad():
            for i in range(n):
                out = net(*batches[i % n_sets])
        return out

    def timed(net, n):
        run(net, max(3, args.warmup))
        torch.cuda.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        run(net, n)
        e1.record(main)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1), dev)

    ms = timed(ours, K)
    ms_ref = timed(ref, max(3, min(K, 10))) / max(3, min(K, 10)) * K if rank == 0 else None
    with torch.no_grad():
        R, t = ours(*batches[0])
        R_ref, t_ref = ref(*batches[0])
    parity = {"against": "the unpatched reference on the same GPU, same weights and inputs (TF32 off)",
              "R_abs_err_max": (R - R_ref).abs().max().item(), "t_abs_err_max": (t - t_ref).abs().max().item(),
              "t_abs_max": t_ref.abs().max().item()}

    # e2e: RGB-D from pinned host memory every step, poses back
    pose_host = [torch.empty((B, 3, 3)).pin_memory(), torch.empty((B, 3)).pin_memory()]

    def e2e_step(i):
        dev_in = tuple(x.to(dev, non_blocking=True) for x in host[i % n_sets])
        with torch.no_grad():
            Ro, to = ours(*dev_in)
        pose_host[0].copy_(Ro, non_blocking=True)
        pose_host[1].copy_(to.reshape(B, 3), non_blocking=True)
    for i in range(3):
        e2e_step(i)
    torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for i in range(K):
        e2e_step(i)
    e1.record(main)
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1), dev)
    torch.cuda.synchronize()
    h2d = sum(x.numel() * x.element_size() for x in host[0])
    return {
        "value": world * B * K / (ms * 1e-3), "ms_per_step": ms / K,
        "config": dict(base_config(wl), flags=" ".join(tracker_flags(wl)),
                       step="forward of LeastSquareTracking in eval mode: the reference's feature encoder (and M-estimator CNN / damping MLP) "
                            "on cuDNN, the solver levels on this repository's CUDA kernels",
                       api="patch_tracker(LeastSquareTracking" + (", fused_forward=True" if (args.workload == "tracker" and not args.level_by_level) else "")
                           + ")(img0, img1, depth0, depth1, K)",
                       l2=l2_note(wl, args.workload, args)),
        "latency_ms": ms / K, "latency_note": "a step is one forward call: its time IS the latency",
        "step_hbm_frac": None, "roofline": None, "parity": parity,
        "e2e": {"value": world * B * K / (ms_e2e * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": B * 12 * 4, "steps": K, "ms_per_step": ms_e2e / K,
                "what": "RGB-D batch in pinned host memory -> device -> forward -> poses in pinned memory, every step"},
        "gpu_launches": None,
        "extra": {"reference_on_this_gpu": {"what": "the UNPATCHED reference (its own PyTorch op chain for the solver levels) on the same GPU, "
                                                     "same weights and inputs", "ms_per_step": ms_ref / K if ms_ref else None,
                                            "pairs_per_s": B * K / (ms_ref * 1e-3) if ms_ref else None,
                                            "speedup": ms_ref / ms if ms_ref else None}},
    }


def run_train(args, wl, rank, world, dev, barrier, max_over_ranks):
    """BASELINE config 4: one data-parallel training step per batch of 64 pairs per GPU (train.py:117-192): the
    reference's LeastSquareTracking with its tr_update0..3 swapped for the CUDA-backed solver (patch_tracker), the
    CUDA pose-pyramid loss, backward, bucketed NCCL all-reduce of the gradients overlapped with the backward, gradient
    clipping and the Adam step."""
    import torch.distributed as dist
    from deep_prob_feature_track_b200 import algorithms as A, criterions
    from deep_prob_feature_track_b200.ddp import FlatBucketReducer, broadcast_parameters
    from baseline import reference as REF
    if not REF.available():
        raise SystemExit("--workload train needs the reference's encoder: baseline/_ref is missing (python baseline/install_reference.py)")
    B, H, W = wl["B"], wl["H"], wl["W"]
    K = args.steps
    main = torch.cuda.current_stream(dev)
    torch.backends.cudnn.benchmark = True
    net = A.patch_tracker(REF.make_tracker(REF.TRAIN_TUM_FLAGS, seed=0)).to(dev).train()
    broadcast_parameters(net)
    reducer = FlatBucketReducer(net.parameters(), n_buckets=4)
    opt = torch.optim.Adam(net.parameters(), lr=1e-5, weight_decay=4e-4)
    batches = [REF.synthetic_rgbd(B, H, W, seed=50 + 10 * rank + i, device=dev) for i in range(2)]
    R_gt = torch.eye(3, device=dev).repeat(B, 1, 1)
    t_gt = torch.tensor([[0.01, -0.005, 0.002]], device=dev).repeat(B, 1)
    parts = {}

    def step(i, timed=False):
        img0, img1, d0, d1, Kc = batches[i % 2]
        reducer.zero_grad()
        if timed:
            marks = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            marks[0].record(main)
        Rs, ts = net(img0, img1, d0, d1, Kc)
        loss = criterions.compute_RT_EPE_loss(Rs, ts, R_gt, t_gt, d0, Kc, invalid=(d0 == d0.min()) | (d0 == d0.max())).mean() * 1e2
        if timed:
            marks[1].record(main)
        loss.backward()
        if timed:
            marks[2].record(main)
        reducer.finish()
        if timed:
            marks[3].record(main)
        torch.nn.utils.clip_grad_norm_(net.parameters(), 5.0)
        opt.step()
        if timed:
            marks[4].record(main)
            torch.cuda.synchronize()
            for name, a, b in (("forward_loss_ms", 0, 1), ("backward_ms", 1, 2), ("allreduce_wait_ms", 2, 3), ("clip_adam_ms", 3, 4)):
                parts.setdefault(name, []).append(marks[a].elapsed_time(marks[b]))
        return loss

    for i in range(max(3, args.warmup)):
        step(i)
    torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for i in range(K):
        loss = step(i)
    e1.record(main)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1), dev)
    loss_value = float(loss.item())
    for i in range(5):
        step(i, timed=True)
    # the collective alone: one all-reduce of the whole flat gradient buffer
    ar_ms = None
    if world > 1:
        for _ in range(3):
            dist.all_reduce(reducer.flat)
        torch.cuda.synchronize()
        ar_ms = event_time_ms(lambda: [dist.all_reduce(reducer.flat) for _ in range(10)], main) / 10
    return {
        "value": world * B * K / (ms * 1e-3), "ms_per_step": ms / K,
        "config": dict(base_config(wl), step="forward (reference FeaturePyramid + pose predictor on cuDNN, CUDA solver 4 levels x 3 iterations) "
                       "+ compute_RT_EPE_loss (CUDA) + backward + bucketed all-reduce + clip_grad_norm + Adam",
                       api="patch_tracker(LeastSquareTracking) + criterions.compute_RT_EPE_loss + ddp.FlatBucketReducer",
                       gradient_bytes=reducer.nbytes, buckets=len(reducer.bounds), collective="NCCL all-reduce" if world > 1 else "none (1 rank)",
                       l2=l2_note(wl, "train", args)),
        "latency_ms": ms / K, "latency_note": "a training step IS the latency",
        "step_hbm_frac": None,
        "roofline": None,
        "parity": {"against": "tests/test_reference_dropin_gpu.py (poses and encoder gradients of this step vs the unpatched reference on CUDA)"},
        "e2e": None,
        "gpu_launches": None,
        "extra": {"step_parts_ms": {k: statistics.median(v) for k, v in parts.items()}, "loss": loss_value,
                  "allreduce_alone_ms": ar_ms, "allreduce_bytes": reducer.nbytes,
                  "note": "allreduce_wait_ms is what the step still waits for after backward: the buckets were started from "
                          "autograd hooks while backward ran"},
    }


def measured_peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except Exception:
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--workload", default="tum", choices=tuple(WORKLOADS))
    ap.add_argument("--batches-per-call", type=int, default=20, help="batches stacked into one solver call (tum)")
    ap.add_argument("--streams", type=int, default=2, help="CUDA streams the calls rotate over")
    ap.add_argument("--no-graphs", action="store_true", help="submit every call launch by launch instead of replaying its CUDA graph")
    ap.add_argument("--frames-per-step", type=int, default=0,
                    help="vga: live frames tracked against the keyframe per step / call (default 16; BASELINE's 1024 pairs over 8 GPUs are 128 per GPU)")
    ap.add_argument("--queue-ctas", type=int, default=0, help="tum: CTAs of the work-queue launch (0 = what the device holds)")
    ap.add_argument("--level-by-level", action="store_true", help="tracker: keep LeastSquareTracking.forward's four module calls")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]
    if args.workload == "vga" and args.frames_per_step > 0:
        wl = dict(wl, B=args.frames_per_step, name=wl["name"].replace("_b16_", f"_b{args.frames_per_step}_"))

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args, wl)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback exists)")
    from deep_prob_feature_track_b200.batched import bind_to_gpu_numa_node
    numa_cpus = bind_to_gpu_numa_node(local_rank)     # before any pinned allocation: first touch on the GPU's node
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from deep_prob_feature_track_b200.sharding import max_over_ranks

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with ClockSampler(local_rank) as clocks:
        body = {"tum": run_tum, "vga": run_vga, "train": run_train, "deepic": run_tracker, "icp": run_tracker,
                "tracker": run_tracker}[args.workload](
            args, wl, rank, world, dev, barrier, max_over_ranks)

    if rank == 0:
        out = {"metric": "frame-pair GN solves/sec", "value": body.pop("value"), "unit": "pairs/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": body.pop("ms_per_step"), "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
        # `config` = what is measured (the reference arm prints the same dict); `how` = how this arm runs it
        full = body.pop("config")
        out["config"] = workload_config(wl, args.workload, args)
        out["how"] = {k: v for k, v in full.items() if k not in out["config"]}
        for k in out["config"]:                       # (the run's own value wins should the two ever disagree)
            if k in full and full[k] != out["config"][k]:
                out["config"][k] = full[k]
        out.update(body)
        out["clocks"] = clocks.summary()
        out["numa"] = {"cpus_bound": numa_cpus}
        if not args.no_cpu_baseline and world == 1:        # (the contract: on rank 0 at N = 1 only)
            out["cpu_baseline"] = cpu_sample(wl, args.workload, 1234)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
