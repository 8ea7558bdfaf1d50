#!/usr/bin/env python
"""Where a kernel's warps spend their time: stall samples and executed instructions per SASS region, from the source
page of an .ncu-rep (captured with --import-source on).  usage: ncu_phases.py file.ncu-rep [instructions_per_bucket]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
step = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
k = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[k], rows[k + 1:]
iS, iW, iE, iT = (hdr.index(n) for n in ("Source", "Warp Stall Sampling (All Samples)", "Instructions Executed", "Thread Instructions Executed"))
tot = sum(int(r[iW]) for r in data) or 1
totE = sum(int(r[iE]) for r in data) or 1
print(f"# {rows[0][1][:100] if len(rows[0]) > 1 else rep}")
print(f"# {len(data)} SASS instructions, {totE} warp instructions executed, {tot} stall samples; buckets of {step} instructions")
print(f"{'sass range':>12s} {'samples':>8s} {'instr':>7s} {'lanes':>6s}  first instruction of the bucket")
for b in range(0, len(data), step):
    rs = data[b:b + step]
    w = sum(int(r[iW]) for r in rs); e = sum(int(r[iE]) for r in rs); t = sum(int(r[iT]) for r in rs)
    if w / tot < 0.004 and e / totE < 0.004:
        continue
    print(f"{b:5d}-{b + len(rs) - 1:<6d} {100 * w / tot:7.1f}% {100 * e / totE:6.1f}% {t / max(e, 1):6.1f}  {rs[0][iS].strip()[:60]}")
