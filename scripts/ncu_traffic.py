#!/usr/bin/env python
"""Per-map DRAM traffic and kernel time from an ncu launch list that carries three metrics per launch
(gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum):
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file launches3.csv python scripts/dev/time_map.py C3 replay
    python scripts/ncu_traffic.py launches3.csv profiles/r02_traffic_c3.json
A map = the launches from one bin_points kernel up to the next.  The LAST map of the run is reported (buffers at their
steady-state sizes).  bench.py quotes the total as roofline.traffic, with this file as its source."""
import collections
import csv
import datetime
import json
import sys


def main(path, out_path):
    lines = [l for l in open(path) if not l.startswith("==")]
    per = collections.OrderedDict()  # launch id -> {name, metrics}
    for row in csv.DictReader(lines):
        d = per.setdefault(row["ID"], {"name": row["Kernel Name"]})
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        m = row["Metric Name"]
        if m.startswith("gpu__time"):
            v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v  # -> us
        else:
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        d[m] = v
    launches = list(per.values())
    starts = [i for i, d in enumerate(launches) if "bin_points" in d["name"]]
    if not starts:
        raise SystemExit("no bin_points launch in the list")
    sel = launches[starts[-1]:]
    agg = collections.OrderedDict()
    for d in sel:
        name = d["name"].split("(")[0].replace("aos::", "").replace("void ", "").replace("<unnamed>::", "")
        a = agg.setdefault(name, {"launches": 0, "us": 0.0, "dram_read": 0.0, "dram_write": 0.0})
        a["launches"] += 1
        a["us"] += d.get("gpu__time_duration.sum", 0.0)
        a["dram_read"] += d.get("dram__bytes_read.sum", 0.0)
        a["dram_write"] += d.get("dram__bytes_write.sum", 0.0)
    tot = {k: sum(a[k] for a in agg.values()) for k in ("launches", "us", "dram_read", "dram_write")}
    out = {"source": path, "captured": datetime.date.today().isoformat(),
           "note": "one C3 map (last of the run), every kernel of aos_map_to_graph; ncu serialises the launches and flushes "
                   "the caches between them, so times and bytes are cold-cache upper bounds",
           "launches": tot["launches"], "kernel_us": round(tot["us"], 1),
           "dram_bytes_read": int(tot["dram_read"]), "dram_bytes_write": int(tot["dram_write"]),
           "dram_bytes": int(tot["dram_read"] + tot["dram_write"]),
           "kernels": {k: {"launches": a["launches"], "us": round(a["us"], 1), "dram_read": int(a["dram_read"]),
                           "dram_write": int(a["dram_write"])}
                       for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"])}}
    json.dump(out, open(out_path, "w"), indent=1)
    print(f"{out_path}: {tot['launches']} launches, {tot['us']:.0f} us, {out['dram_bytes'] / 1e9:.3f} GB DRAM traffic per map")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
