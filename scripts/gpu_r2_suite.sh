#!/bin/bash
# 1 GPU: the whole -m gpu suite, then the default bench and the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/suite.txt
timeout 1400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -6 gpurun_out/bench_default.err
python scripts/summ.py default < gpurun_out/bench_default.json | tee -a gpurun_out/suite.txt
