#!/bin/bash
for cfg in "--unroll 2" "--unroll 4"; do
python bench.py $cfg --no-cpu --no-e2e --steps 400 --warmup 100 2>/dev/null | python scripts/summ.py "$cfg minctas8"
done
