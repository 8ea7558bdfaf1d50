#!/usr/bin/env python
"""Top SASS instructions by stall samples from an .ncu-rep source page: ncu_source.py file.ncu-rep [N]"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
rows = list(csv.reader(lines[1:]))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[1:]
tot = sum(int(r[ix["# Samples"]]) for r in data if r[ix["# Samples"]].isdigit())
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print("total samples", tot, "instructions", len(data))
stalls = ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_barrier", "stall_branch_resolving",
          "stall_not_selected", "stall_selected", "stall_lg", "stall_mio", "stall_dispatch", "stall_no_inst"]
agg = {s: sum(int(r[ix[s]]) for r in data if r[ix[s]].isdigit()) for s in stalls}
print({k: round(v / max(tot, 1), 3) for k, v in agg.items()})
top = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:n]
for i in sorted(top):
    r = data[i]
    st = {s.replace("stall_", ""): int(r[ix[s]]) for s in stalls if r[ix[s]].isdigit() and int(r[ix[s]]) > 0}
    print(f"{i:4d} {int(r[ix['# Samples']]):6d} {100*int(r[ix['# Samples']])/tot:5.1f}% exec={r[ix['Instructions Executed']]:>9s} thr={r[ix['Avg. Threads Executed']]:>4s} {r[ix['Source']].strip()[:70]:70s} {st}")
