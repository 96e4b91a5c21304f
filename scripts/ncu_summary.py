#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of the step)
and, optionally, pull the DRAM traffic of one `ncu --set full` report.  Usage:
    python scripts/ncu_summary.py gpurun_out/launches.csv [gpurun_out/prof.ncu-rep] > profiles/rNN_....txt
"""
import collections
import csv
import subprocess
import sys


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        a = agg.setdefault(row["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot:.1f} us total (cold-cache, serialised: compare shares)")
    print(f"{'us':>12} {'share':>6} {'n':>5}  kernel")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:12.1f} {100 * t / tot:5.1f}% {n:5d}  {k[:110]}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        print("# could not read", path)
        return
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
            "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
            "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct"]
    print(f"# {path}")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        for w in want:
            if w in d:
                print(f"{w:75s} {d[w]:>20s} {units[hdr.index(w)]}")
        print()


if __name__ == "__main__":
    launches(sys.argv[1])
    for p in sys.argv[2:]:
        print()
        full(p)
