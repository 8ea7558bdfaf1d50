#!/bin/bash
# quick single-GPU sweep used during development (run under gpurun)
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for cfg in "--tpa 4 --unroll 2" "--tpa 4 --unroll 4" "--tpa 2 --unroll 2" "--tpa 2 --unroll 4" "--tpa 2 --unroll 2 --force-grid 1184" "--tpa 2 --unroll 2 --force-grid 4736" "--tpa 1 --unroll 4"; do
  python bench.py $cfg --no-cpu --no-e2e --steps 500 --warmup 100 2>/dev/null | python scripts/summ.py "$cfg"
done
