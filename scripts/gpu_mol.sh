#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_scale.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -15
for w in butane water; do
timeout 900 python bench.py --workload $w --steps 300 --warmup 50 2>gpurun_out/b_$w.err > gpurun_out/b_$w.json; tail -3 gpurun_out/b_$w.err
python - <<PY
import json
d=json.load(open("gpurun_out/b_$w.json"))
print("$w", "value %.3e" % d["value"], "ms/step %.3f" % d["ms_per_step"], {k:(round(v["total_ms"]/max(v["launches"],1),4), v["launches"]) for k,v in d["kernel_ms"].items()}, "e2e %.3e" % d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], "pairs/atom", d["config"]["half_pairs_per_atom"], "rebuilds", d["config"]["list_rebuilds_in_timed_region"])
PY
done
