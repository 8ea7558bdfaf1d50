#!/bin/bash
for cfg in "--tpa 1 --force-grid 2368" "--tpa 1 --force-grid 3552" "--tpa 1 --force-grid 4736" "--tpa 1 --force-grid 9472" "--tpa 1 --force-grid 18944"; do
  python bench.py $cfg --no-cpu --no-e2e --steps 500 --warmup 100 2>/dev/null | python scripts/summ.py "$cfg"
done
ncu --set full --clock-control none --import-source on -k regex:k_lj_list -s 6 -c 1 -o gpurun_out/prof_force python bench.py --tpa 1 --steps 8 --warmup 8 --no-cpu --no-e2e > gpurun_out/p1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_build_tile -s 1 -c 1 -o gpurun_out/prof_build python bench.py --tpa 1 --steps 8 --warmup 8 --no-cpu --no-e2e > gpurun_out/p2.log 2>&1
