"""Round 2 probe: S streams x calls of G batches each (stream_probe_items.txt is the log of the version that also
tried queue workers that leave after 1 / 2 / 4 items: no gain, option removed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from deep_prob_feature_track_b200 import algorithms as A
from deep_prob_feature_track_b200.synthetic import make_frame_pairs

dev = torch.device("cuda:0")
B, C, H, W = 64, 8, 120, 160
gmax = 8
parts = [make_frame_pairs(B, C, H, W, seed=1234 + g, n_levels=4) for g in range(gmax)]
levels = [{k: torch.cat([p["levels"][l][k] for p in parts]).to(dev) for k in parts[0]["levels"][l]} for l in range(4)]
pose = (torch.cat([p["R0"] for p in parts]).to(dev), torch.cat([p["t0"] for p in parts]).to(dev))
sets = []
for s in range(4):
    sets.append(([{k: torch.roll(v, s * 3, 0).contiguous() for k, v in lv.items()} for lv in levels], pose))


def sub(levels, pose, n):
    return [{k: v[:n] for k, v in lv.items()} for lv in levels], (pose[0][:n], pose[1][:n])


def run(G, S, n_calls, **kw):
    streams = [torch.cuda.Stream(device=dev) for _ in range(S)]
    data = [sub(*s, B * G) for s in sets]
    def go(n):
        for i in range(n):
            with torch.cuda.stream(streams[i % S]):
                A.uic_solve(*data[i % 4], iters=3, remove_tru_sigma=True, group=B, queue=True, **kw)
    go(2 * S)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for st in streams:
        st.wait_event(e0)
    go(n_calls)
    for st in streams:
        ev = torch.cuda.Event(); ev.record(st); torch.cuda.current_stream().wait_event(ev)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (n_calls * G)


for G in (2, 4, 8):
    for S in (1, 2, 3, 4):
        for tr in (30, 40):
            t = run(G, S, max(8, 48 // G), tile_rows=[0, 0, 0, tr])
            print(f"G={G} streams={S} tile_rows={tr}: {t:7.1f} us per batch", flush=True)
