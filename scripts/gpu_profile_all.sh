#!/bin/bash
# Round profile set (1 GPU): launch list of the bench command + full captures of the three dominant kernels.
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 10 --no-cpu --no-e2e > gpurun_out/launches.log 2>&1
B="python bench.py --steps 10 --warmup 100 --no-cpu --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:k_lj_list -s 100 -c 1 -f -o gpurun_out/prof_k_lj_list $B > gpurun_out/prof_k_lj_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_build_tile2 -s 15 -c 1 -f -o gpurun_out/prof_k_build_tile2 $B > gpurun_out/prof_k_build_tile2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 100 -c 1 -f -o gpurun_out/prof_k_integrate $B > gpurun_out/prof_k_integrate.log 2>&1
ls -la gpurun_out/ | tail -12
