#!/bin/bash
# Round 2, tile kernels on hardware (1 GPU): parity tests of the tile path, bench A/B, full ncu captures.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lj.py tests/test_gpu_zzz_options.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2_tile.txt
for o in "row_sched=0"; do
  SEPGPU_OPTS="$o" timeout 400 python bench.py --steps 600 --warmup 200 --no-cpu --no-e2e 2>gpurun_out/r2_tile_b.err >gpurun_out/r2_tile_b_$o.json
  python scripts/summ.py "$o" < gpurun_out/r2_tile_b_$o.json | tee -a gpurun_out/r2_tile.txt
  tail -3 gpurun_out/r2_tile_b.err
done
B="python bench.py --steps 10 --warmup 100 --no-cpu --no-e2e"
SEPGPU_OPTS="${PROF_OPTS:-row_sched=0}" ncu --set full --clock-control none --import-source on -k regex:k_lj_tile -s 100 -c 1 -f -o gpurun_out/prof_k_lj_tile $B > gpurun_out/prof_k_lj_tile.log 2>&1
SEPGPU_OPTS="${PROF_OPTS:-row_sched=0}" ncu --set full --clock-control none --import-source on -k regex:k_build_tile -s 15 -c 1 -f -o gpurun_out/prof_k_build_tile $B > gpurun_out/prof_k_build_tile.log 2>&1
ls -la gpurun_out/ | tail -4
