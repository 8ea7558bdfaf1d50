#!/bin/bash
# water (C3): launch list of one rebuild cycle + full captures of its builder and Coulomb kernel
mkdir -p gpurun_out
B="python bench.py --workload water --steps 20 --warmup 20 --no-cpu --no-e2e --no-other"
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 120 --csv --log-file gpurun_out/water_launches.csv $B > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/water_launches.csv")) if len(r) > 5 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    k = r[4].split("(")[0][:60]; agg[k][0] += 1; agg[k][1] += float(r[-1]) / 1e3
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]): print("%-60s %4d launches %9.1f us total %8.1f us each" % (k, n, t, t / n))
PY
ncu --set full --clock-control none --import-source on -k regex:k_build_tile -s 2 -c 1 -f -o gpurun_out/prof_water_build $B > gpurun_out/prof_water_build.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_coul -s 30 -c 1 -f -o gpurun_out/prof_water_coul $B > gpurun_out/prof_water_coul.log 2>&1
ls -la gpurun_out | tail -3
