#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel launches, total and share.
usage: launch_summary.py launches.csv [first_launch_id] [exclude_regex]
(first_launch_id skips set-up launches; exclude_regex drops probes that run outside the timed region)"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
excl = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
agg = OrderedDict()
for r in rows:
    if int(r[0]) < skip:
        continue
    name = re.sub(r"\(.*", "", r[4])
    name = re.sub(r"^void ", "", name)
    if excl and excl.search(name):
        continue
    a = agg.setdefault(name, [0, 0.0, r[7], r[8]])
    a[0] += 1
    a[1] += float(r[14]) / 1e3
tot = sum(a[1] for a in agg.values())
print(f"# launches {sum(a[0] for a in agg.values())} (ids >= {skip}), total kernel time {tot / 1e3:.3f} ms (cold-cache, serialised under ncu)")
print(f"{'kernel':44s} {'launches':>8s} {'total_us':>10s} {'avg_us':>9s} {'share':>7s}  block / grid")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:44]:44s} {a[0]:8d} {a[1]:10.1f} {a[1] / a[0]:9.1f} {100 * a[1] / tot:6.1f}%  {a[2]} / {a[3]}")
