#!/bin/bash
# 1 GPU, quick: parity tests of the list/force path, then the kernel-only bench (optionally with SEPGPU_OPTS variants)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lj.py tests/test_gpu_zzz_options.py tests/test_golden.py tests/test_gpu_more.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/quick.txt
for o in ${QUICK_OPTS:-none}; do
  [ "$o" = none ] && o=""
  SEPGPU_OPTS="$o" timeout 400 python bench.py --steps 600 --warmup 200 --no-cpu --no-e2e ${QUICK_ARGS} 2>gpurun_out/quick_b.err >gpurun_out/quick_b_$o.json
  python scripts/summ.py "opts=$o" < gpurun_out/quick_b_$o.json | tee -a gpurun_out/quick.txt
  tail -3 gpurun_out/quick_b.err
done
