#!/bin/bash
python bench.py --no-cpu --steps 300 --warmup 50 2>gpurun_out/e.err | python scripts/summ.py "default"
grep "e2e breakdown" gpurun_out/e.err
