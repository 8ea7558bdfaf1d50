#!/bin/bash
# 1-GPU: full GPU test suite, then a default bench run
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py 2>gpurun_out/b1.err > gpurun_out/b1.json; python scripts/summ.py "N=1" < gpurun_out/b1.json
tail -5 gpurun_out/b1.err
