#!/bin/bash
# quick single-GPU sweep used during development (run under gpurun)
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for t in 4 2 8 1; do
  python bench.py --tpa $t --no-cpu --no-e2e --steps 500 --warmup 100 2>/dev/null | python scripts/summ.py "tpa=$t"
done
