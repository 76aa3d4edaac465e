"""Turn ncu outputs brought back in gpurun_out/ into the text summaries committed under profiles/.

  python profiles/summarize.py launches gpurun_out/launches.csv          > profiles/rNN_launches.txt
  python profiles/summarize.py kernel   gpurun_out/prof.ncu-rep [rows]   > profiles/rNN_kernel.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ["Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ik, iv, ig = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    agg = collections.defaultdict(list)
    for r in rows[1:]:
        agg[(r[ik][:70], r[ig])].append(float(r[iv].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k[0]:70s} grid={k[1]:16s} n={len(v):4d} avg={sum(v) / len(v) / 1e3:9.1f} us share={sum(v) / tot * 100:5.1f}%")


def kernel(path, rows_per_launch=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("## " + r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                print(f"{k:82s} {units[hdr.index(k)]:16s} {r[hdr.index(k)]}")
        st = [(hdr[i], float(r[i].replace(",", "") or 0)) for i in range(len(hdr))
              if hdr[i].startswith("smsp__pcsamp_warps_issue_stalled") and not hdr[i].endswith("not_issued")]
        tot = sum(v for _, v in st) or 1
        print("stall samples: " + ", ".join(f"{k.split('stalled_')[1]} {v / tot * 100:.0f}%" for k, v in
                                            sorted(st, key=lambda kv: -kv[1])[:7]))
    sass = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                          capture_output=True, text=True).stdout
    hdr, seen, opc, samp, tot = None, 0, collections.Counter(), collections.Counter(), 0
    for r in csv.reader(sass.splitlines()):
        if not r:
            continue
        if r[0] == "Kernel Name":
            seen += 1
            if seen > 1:
                break
            continue
        if r[0] == "Address":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        toks = r[1].split()
        if not toks:
            continue
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        ie = int(r[hdr.index("Instructions Executed")])
        opc[op] += ie
        samp[op] += int(r[hdr.index("# Samples")])
        tot += ie
    print("\n## opcode mix of the first captured launch (share of warp instructions, stall samples)")
    for k, v in opc.most_common(16):
        extra = f" {v / tot * float(rows_per_launch):8.1f}/warp-row" if rows_per_launch else ""
        print(f"{k:8s} {v / tot * 100:5.1f}%  samples {samp[k]:6d}{extra}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        kernel(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
