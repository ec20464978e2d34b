#!/usr/bin/env python3
"""Per-kernel totals and shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
Usage: launch_summary.py launches.csv > profiles/rN_launches_summary.txt"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split("(")[0]
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
    agg.setdefault(name, []).append(v)
tot = sum(sum(v) for v in agg.values())
print("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 3 "
      "--warmup 3 --no-cpu-baseline --e2e-steps 1")
print("# cold-cache, serialised launches: compare SHARES with bench.py's phase_ms, not absolutes")
print("%-60s %8s %12s %7s" % ("kernel", "launches", "total_us", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-60s %8d %12.1f %6.1f%%" % (k, len(v), sum(v), 100 * sum(v) / tot))
