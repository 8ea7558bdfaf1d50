#!/bin/bash
# 1 GPU: the whole -m gpu suite, then the default bench and the reference arm
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/suite.txt
timeout 900 python bench.py 2>gpurun_out/bench_default.err >gpurun_out/bench_default.json; python scripts/summ.py "default" < gpurun_out/bench_default.json
tail -4 gpurun_out/bench_default.err
