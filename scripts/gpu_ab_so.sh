#!/bin/bash
# A/B of alternative builds of libsep.so (SEPLIB_SO=path), 1 GPU.  usage: gpu_ab_so.sh lib1.so lib2.so ...
mkdir -p gpurun_out
: > gpurun_out/ab_so.txt
for so in "$@"; do
  SEPLIB_SO=$PWD/$so timeout 400 python bench.py --steps 600 --warmup 200 --no-cpu --no-e2e 2>gpurun_out/ab_so.err >gpurun_out/ab_so.json
  python scripts/summ.py "$so" < gpurun_out/ab_so.json | tee -a gpurun_out/ab_so.txt
  tail -2 gpurun_out/ab_so.err
done
