#!/usr/bin/env python
"""Summarise bench.py JSON lines read from stdin (one per line): label value ms/step kernel means."""
import json
import sys

label = sys.argv[1] if len(sys.argv) > 1 else ""
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    km = {k: round(v["total_ms"] / max(v["launches"], 1), 4) for k, v in d.get("kernel_ms", {}).items()}
    r = d.get("roofline") or {}
    e = d.get("e2e") or {}
    print(label, "value %.3e ms/step %.4f" % (d["value"], d["ms_per_step"]), "kernels(ms/launch)", km,
          "rebuilds", d["config"].get("list_rebuilds_in_timed_region"),
          "hbm_frac %.3f" % r.get("frac", 0), "fp64_frac %.3f" % ((r.get("fp64") or {}).get("frac") or 0),
          "e2e %.3e" % e.get("value", 0) if e else "")
