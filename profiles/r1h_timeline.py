"""Per-warp timeline of the finest-level launch of the staged kernel (library built with -DDPFT_DEBUG_STAMPS):
when each warp starts and ends its tile walk, on which SM, how many rows.  Prints a summary."""
import ctypes
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deep_prob_feature_track_b200 import _lib, algorithms as A
from deep_prob_feature_track_b200.synthetic import levels_to, make_frame_pairs

B = 64
data = make_frame_pairs(B, 8, 120, 160, seed=1234, n_levels=4)
lv = levels_to(data["levels"], "cuda:0")
pose = (data["R0"].cuda(), data["t0"].cuda())
NW = 8192
for rep in range(3):
    res = A.uic_solve(lv, pose, iters=3, remove_tru_sigma=True, timed=True)
    torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (4 * NW))()
L = _lib.lib()
L.dpft_debug_read_timeline(buf, NW)
st = (ctypes.c_ulonglong * 16)()
L.dpft_debug_read_stamps(st)
def unpack(i):
    v = buf[4 * i + 3]
    return (buf[4 * i] & 0xffffffff, buf[4 * i + 1], buf[4 * i + 2], v & 0xffff, (v >> 16) & 1, (v >> 24) & 0xff, (v >> 32) & 0xffff, v >> 48, (buf[4 * i] >> 32) & 0xff, buf[4 * i] >> 40)


recs = [unpack(i) for i in range(NW) if buf[4 * i + 1]]   # smid, t_start, t_end, rows, two sub-tiles, seg, y0, direct rows
t0 = min(r[1] for r in recs)
print("launch_ms", round(res.launch_ms[-1] * 1e3, 1), "warps", len(recs), "stamps(us rel. first start):",
      " ".join(f"{((st[i] - t0) / 1e3):.1f}" for i in range(8)))
starts = sorted((r[1] - t0) / 1e3 for r in recs)
ends = sorted((r[2] - t0) / 1e3 for r in recs)
q = lambda v, f: v[min(len(v) - 1, int(f * len(v)))]
print("start us: min %.1f p50 %.1f p90 %.1f max %.1f" % (starts[0], q(starts, .5), q(starts, .9), starts[-1]))
print("end   us: min %.1f p10 %.1f p50 %.1f p90 %.1f max %.1f" % (ends[0], q(ends, .1), q(ends, .5), q(ends, .9), ends[-1]))
per_sm = defaultdict(list)
for r in recs:
    per_sm[r[0]].append(r)
by_n = defaultdict(list)
for sm, rs in per_sm.items():
    by_n[len(rs)].append((max(x[2] for x in rs) - t0) / 1e3)
for n, v in sorted(by_n.items()):
    print(f"SMs with {n} warps: {len(v)}  last end us: min {min(v):.1f} mean {sum(v) / len(v):.1f} max {max(v):.1f}")
by_rows = defaultdict(list)
for r in recs:
    if r[3]:
        by_rows[(len(per_sm[r[0]]), r[3], r[4])].append((r[2] - r[1]) / 1e3 / r[3])
print("us per row by (warps on the SM, rows, two sub-tiles):")
for k, v in sorted(by_rows.items()):
    print("  ", k, "n=%d" % len(v), "mean %.2f min %.2f max %.2f" % (sum(v) / len(v), min(v), max(v)))
by_seg = defaultdict(list)
for r in recs:
    if r[3]:
        by_seg[r[5]].append(((r[2] - r[1]) / 1e3 / r[3], r[7] / r[3]))
print("first segment of the warp: us per row, share of rows on the direct-load body")
for k, v in sorted(by_seg.items()):
    print("  seg", k, "n=%d" % len(v), "us/row mean %.2f max %.2f   direct share mean %.3f max %.3f" % (
        sum(x[0] for x in v) / len(v), max(x[0] for x in v), sum(x[1] for x in v) / len(v), max(x[1] for x in v)))
slow = sorted(recs, key=lambda r: -(r[2] - r[1]) / max(r[3], 1))[:12]
print("slowest warps: (us/row, rows, seg, y0, direct rows, warps on SM, end us, ring restarts, non-resident lanes)")
for r in slow:
    print("  ", round((r[2] - r[1]) / 1e3 / max(r[3], 1), 2), r[3], r[5], r[6], r[7], len(per_sm[r[0]]), round((r[2] - t0) / 1e3, 1), r[8], r[9])
print("all warps: ring restarts total", sum(r[8] for r in recs), "non-resident lanes total", sum(r[9] for r in recs), "direct rows total", sum(r[7] for r in recs))
# does the direct share explain the time per row?
xs = [r[7] / r[3] for r in recs if r[3]]
ys = [(r[2] - r[1]) / 1e3 / r[3] for r in recs if r[3]]
mx, my = sum(xs) / len(xs), sum(ys) / len(ys)
cov = sum((a - mx) * (b - my) for a, b in zip(xs, ys)); vx = sum((a - mx) ** 2 for a in xs); vy = sum((b - my) ** 2 for b in ys)
print("corr(direct share, us/row) = %.3f; mean direct share %.3f" % (cov / max((vx * vy) ** 0.5, 1e-30), mx))

# phase clocks per warp (cycles per row): geometry + ring logic | waiting for the ring | rest of the row
buf2 = (ctypes.c_ulonglong * (4 * NW))()
L.dpft_debug_read_phases(buf2, NW)
idx = [i for i in range(NW) if buf[4 * i + 1]]
rows_of = {i: unpack(i) for i in idx}
def phase(i):
    r = rows_of[i]
    n = max(r[3], 1)
    return (buf2[4 * i] / n, buf2[4 * i + 1] / n, buf2[4 * i + 2] / n, (buf2[4 * i + 3] & 0xffffffff) / n, (buf2[4 * i + 3] >> 32) / n,
            (r[2] - r[1]) / 1e3 / n)
ph = sorted((phase(i) for i in idx if rows_of[i][3]), key=lambda t: t[5])
def avg(rows, k): return sum(t[k] for t in rows) / len(rows)
n = len(ph)
for name, grp in (("fastest 10%", ph[: n // 10]), ("middle 80%", ph[n // 10: -n // 10]), ("slowest 10%", ph[-n // 10:]), ("slowest 1%", ph[-n // 100:])):
    print(f"{name:12s} us/row {avg(grp, 5):.2f}  cycles/row front {avg(grp, 0):.0f} wait {avg(grp, 1):.0f} body {avg(grp, 2):.0f}  staged rows/row {avg(grp, 3):.2f}  tru branch/row {avg(grp, 4):.2f}")
