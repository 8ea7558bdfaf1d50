#!/bin/bash
# Round-2 evidence for profiles/: launch list of the bench step, full captures of the three step kernels.
mkdir -p gpurun_out
B="python bench.py --steps 30 --warmup 100 --equilibrate 0 --no-cpu --no-e2e --no-other"
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 300 --csv --log-file gpurun_out/r02_launches_bench_n1.csv $B > /dev/null 2>&1
for k in k_lj_tile k_build_tile k_integrate; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 40 -c 1 -f -o gpurun_out/r02_$k $B > gpurun_out/r02_$k.log 2>&1
done
ls -la gpurun_out | grep r02_
