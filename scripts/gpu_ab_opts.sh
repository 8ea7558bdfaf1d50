#!/bin/bash
# A/B of SEPGPU_OPTS settings, 1 GPU.  usage: gpu_ab_opts.sh workload steps warmup "opts1" "opts2" ...
mkdir -p gpurun_out
wl=$1; st=$2; wu=$3; shift 3
for o in "$@"; do
  SEPGPU_OPTS="$o" timeout 600 python bench.py --workload $wl --steps $st --warmup $wu --no-cpu --no-e2e --no-other 2>gpurun_out/ab.err >gpurun_out/ab.json
  python scripts/summ.py "$wl [$o]" < gpurun_out/ab.json | tee -a gpurun_out/ab_opts.txt
  tail -2 gpurun_out/ab.err
done
