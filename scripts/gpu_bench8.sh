#!/bin/bash
# 8-GPU: the C4 bench (8 M atoms) -- bench.py itself runs the decomposed-vs-single check before its timed region.
# Everything under a tight timeout.
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 300 --warmup 100 --no-cpu 2>gpurun_out/b8.err > gpurun_out/b8.json; python scripts/summ.py "N=8 dd" < gpurun_out/b8.json
grep -v "^W0\|OMP_NUM\|^\*\*\*\|^$" gpurun_out/b8.err | tail -4
