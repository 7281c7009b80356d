#!/usr/bin/env python
"""Summarises an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`) of bench.py:
picks ONE training step (the launches between two consecutive sgd_multi_kernel launches) and prints per-kernel
counts, summed durations and shares.  Usage: python tools/summarize_launches.py launches.csv [step_index]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10 and r[0].isdigit()]
    names = [re.sub(r"\(.*", "", r[4]).replace("b2::", "").replace("void ", "") for r in rows]
    durs = [float(r[-1]) / 1e3 for r in rows]
    ends = [i for i, n in enumerate(names) if n.startswith("sgd_multi_kernel")]
    if len(ends) < 2:
        raise SystemExit("fewer than two optimiser launches in %s" % path)
    k = int(sys.argv[2]) if len(sys.argv) > 2 else len(ends) // 2
    lo, hi = ends[k - 1] + 1, ends[k] + 1
    agg = collections.OrderedDict()
    for n, t in zip(names[lo:hi], durs[lo:hi]):
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(a[1] for a in agg.values())
    print("# one training step = launches %d..%d of %d (%d launches), %.1f us summed (cold-cache, serialised)"
          % (lo, hi - 1, len(rows), hi - lo, tot))
    print("%-48s %5s %10s %7s" % ("kernel", "count", "sum_us", "share"))
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-48s %5d %10.1f %6.1f%%" % (n[:48], c, t, 100 * t / tot))


if __name__ == "__main__":
    main()
