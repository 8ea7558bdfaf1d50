#!/bin/bash
# 8-GPU: decomposition parity check, then the C4 bench (8 M atoms)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tests/dd_check.py 2>&1 | grep "dd_check\|differ\|rror" | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 300 --warmup 100 2>gpurun_out/b8.err > gpurun_out/b8.json; python scripts/summ.py "N=8 dd" < gpurun_out/b8.json
grep -v "^W0\|OMP_NUM\|^\*\*\*\|^$" gpurun_out/b8.err | tail -3
