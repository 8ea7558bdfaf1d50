#!/bin/bash
# 1 GPU: the whole -m gpu suite (no -x: every failure is listed), then the default bench
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/suite.txt
