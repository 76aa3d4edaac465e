"""Round 2e probe: the replication check of the work-queue levels on the library's side stream (default) against the whole
check up front on the caller's stream (sigma_detect = 2) and against no check (sigma_detect = 1, C-map routines).
Needs r2e_side_stream_check_experiment.patch applied (sigma_detect = 2 is that patch's knob).
Usage: python profiles/r2/side_check_probe.py [batches ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.batched import BatchedSolver
from deep_prob_feature_track_b200.synthetic import make_frame_pairs

dev = torch.device("cuda:0")
B, C, H, W = 64, 8, 120, 160
groups = [int(x) for x in sys.argv[1:]] or [4, 8, 20]
gmax = max(groups)
sets = []
for s in range(2):
    parts = [make_frame_pairs(B, C, H, W, seed=1234 + 17 * s + g, n_levels=4) for g in range(gmax)]
    levels = [{k: torch.cat([p["levels"][l][k] for p in parts]).to(dev) for k in parts[0]["levels"][l]} for l in range(4)]
    for lv in levels:     # the bench's format: sigma repeated to C channels
        lv["s0"] = lv["s0"].expand(-1, C, -1, -1).contiguous(); lv["s1"] = lv["s1"].expand(-1, C, -1, -1).contiguous()
    pose = (torch.cat([p["R0"] for p in parts]).to(dev), torch.cat([p["t0"] for p in parts]).to(dev))
    sets.append((levels, pose))


def sub(levels, pose, n):
    return [{k: v[:n] for k, v in lv.items()} for lv in levels], (pose[0][:n], pose[1][:n])


def timeit(fn, n=20, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3   # us


for G in groups:
    data = [sub(*s, B * G) for s in sets]
    ref = None
    for name, tun in (("check on the side stream", None), ("check up front        ", dict(sigma_detect=2)), ("no check (C-map)      ", dict(sigma_detect=1))):
        kw = dict(iters=3, remove_tru_sigma=True, group=B, tuning=tun)
        t = timeit(lambda i: A.uic_solve(*data[i % 2], **kw), n=max(6, 80 // G))
        r = A.uic_solve(*data[0], **kw)
        torch.cuda.synchronize()
        ref = ref or r
        line = f"G={G:2d} eager, {name}: {t / G:7.1f} us per batch ({t:7.1f} us per call); equal to the first: {torch.equal(r.pose_hist, ref.pose_hist) and torch.equal(r.sys_hist, ref.sys_hist)}"
        if tun is None or tun.get("sigma_detect") == 2:
            solver = BatchedSolver(B, iters=3, remove_tru_sigma=True, streams=1, device=dev, graphs=True, tuning=tun)
            for i in range(4):
                res = solver.submit(*data[i % 2])
            solver.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = max(6, 80 // G)
            main = torch.cuda.current_stream()
            e0.record(main); solver.wait_for(e0)
            for i in range(n):
                res = solver.submit(*data[i % 2])
            solver.join(main); e1.record(main)
            torch.cuda.synchronize()
            again = solver.submit(*data[0]); solver.synchronize()
            line += f";  graph replay {e0.elapsed_time(e1) / n * 1e3:7.1f} us per call ({solver.replays} replays), equal: {torch.equal(again.pose_hist, ref.pose_hist)}"
        print(line, flush=True)
    # independent sigma channels: the flags say 'not replicated', the C-map twins run -- same results as without the check
    lv2 = [dict(lv) for lv in data[0][0]]
    lv2[-1]["s1"] = lv2[-1]["s1"].clone(); lv2[-1]["s1"][5, 3, 7, 9] *= 1.5
    a = A.uic_solve(lv2, data[0][1], iters=3, remove_tru_sigma=True, group=B)
    b = A.uic_solve(lv2, data[0][1], iters=3, remove_tru_sigma=True, group=B, tuning=dict(sigma_detect=1))
    c = A.uic_solve(lv2, data[0][1], iters=3, remove_tru_sigma=True, group=B, tuning=dict(sigma_detect=2))
    torch.cuda.synchronize()
    print(f"G={G:2d} one sigma1 element of the finest level changed: finest level equal to the unchecked solve "
          f"{torch.equal(a.pose_hist[-3:], b.pose_hist[-3:])} / {torch.equal(c.pose_hist[-3:], b.pose_hist[-3:])}; "
          f"max |pose diff| {(a.pose_hist - b.pose_hist).abs().max().item():.1e}", flush=True)
