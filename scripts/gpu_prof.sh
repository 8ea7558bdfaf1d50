#!/bin/bash
# ncu captures of the two dominant kernels (run under gpurun, 1 GPU)
ncu --set full --clock-control none --import-source on -k regex:k_lj_list -s 6 -c 1 -o gpurun_out/prof_force python bench.py --steps 8 --warmup 8 --no-cpu --no-e2e > gpurun_out/p1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_build_tile -s 1 -c 1 -o gpurun_out/prof_build python bench.py --steps 8 --warmup 8 --no-cpu --no-e2e > gpurun_out/p2.log 2>&1
ls -la gpurun_out
