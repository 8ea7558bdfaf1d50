#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 100 --equilibrate 0 --no-cpu --no-e2e --no-other"
ncu --set full --clock-control none --import-source on -k regex:k_build_tile -s 15 -c 1 -f -o gpurun_out/prof_k_build_tile $B > gpurun_out/prof_k_build_tile.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_cell|k_scan|k_build" -s 60 -c 40 --csv --log-file gpurun_out/build_launches.csv $B > /dev/null 2>&1
tail -45 gpurun_out/build_launches.csv | cut -d, -f5,11- | head -45
