#!/usr/bin/env python
"""Per-kernel summary of an ncu metrics CSV (tools/ncu_step_metrics.sh) over ONE training step (the launches between
two sgd_multi_kernel launches): count, time, DRAM bytes, achieved DRAM GB/s and % of the measured HBM peak
(MEASURED_PEAKS.json), time-weighted tensor-pipe utilisation.  Usage: summarize_step_metrics.py file.csv"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    peak = 6540.8
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
    hdr = next(r for r in rows if r[0] == "ID")
    data = [r for r in rows if r[0].isdigit()]
    ki, mi, ui, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    launches = collections.OrderedDict()
    for r in data:
        d = launches.setdefault(int(r[0]), {"name": re.sub(r"\(.*", "", r[ki]).replace("b2::", "").replace("void ", "")})
        v = float(r[vi].replace(",", "")) if r[vi] not in ("", "n/a") else 0.0
        u = r[ui]
        if u in ("Mbyte",): v *= 1e6
        elif u in ("Kbyte",): v *= 1e3
        elif u in ("Gbyte",): v *= 1e9
        elif u in ("us", "usecond"): v *= 1e-6
        elif u in ("ns", "nsecond"): v *= 1e-9
        elif u in ("ms", "msecond"): v *= 1e-3
        d[r[mi]] = v
    ids = sorted(launches)
    ends = [i for i in ids if launches[i]["name"].startswith("sgd_multi_kernel")]
    if len(ends) < 2:
        raise SystemExit("fewer than two optimiser launches")
    k = len(ends) // 2
    step = [launches[i] for i in ids if ends[k - 1] < i <= ends[k]]
    agg = collections.OrderedDict()
    for d in step:
        a = agg.setdefault(d["name"], {"n": 0, "t": 0.0, "rd": 0.0, "wr": 0.0, "tensor_t": 0.0})
        t = d.get("gpu__time_duration.sum", 0.0)
        a["n"] += 1
        a["t"] += t
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
        a["tensor_t"] += t * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
    tot = sum(a["t"] for a in agg.values())
    print("# one training step: %d launches, %.1f us summed (ncu: cold cache, serialised); HBM peak %.1f GB/s (measured)"
          % (len(step), tot * 1e6, peak))
    print("%-44s %5s %9s %6s %10s %10s %9s %7s %8s" % ("kernel", "count", "sum_us", "share", "dram_rd_MB", "dram_wr_MB",
                                                       "GB/s", "%peak", "tensor%"))
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
        gbs = (a["rd"] + a["wr"]) / a["t"] / 1e9 if a["t"] > 0 else 0.0
        print("%-44s %5d %9.1f %5.1f%% %10.1f %10.1f %9.0f %6.1f%% %7.1f%%"
              % (n[:44], a["n"], a["t"] * 1e6, 100 * a["t"] / tot, a["rd"] / 1e6, a["wr"] / 1e6, gbs, 100 * gbs / peak,
                 a["tensor_t"] / a["t"] if a["t"] > 0 else 0.0))


if __name__ == "__main__":
    main()
