#!/bin/bash
ncu --set full --clock-control none --import-source on -k regex:k_build_tile2 -s 1 -c 1 -o gpurun_out/prof_build python bench.py --steps 8 --warmup 8 --no-cpu --no-e2e > gpurun_out/p2.log 2>&1
