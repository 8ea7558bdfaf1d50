#!/bin/bash
# 8-GPU: decomposition parity check (uneven layer split), then the C4 bench (8 M atoms).  Everything under a tight timeout.
DD_NCELL=48 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tests/dd_check.py 2>&1 | grep "dd_check\|differ\|rror" | tail -3
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 300 --warmup 100 2>gpurun_out/b8.err > gpurun_out/b8.json; python scripts/summ.py "N=8 dd" < gpurun_out/b8.json
grep -v "^W0\|OMP_NUM\|^\*\*\*\|^$" gpurun_out/b8.err | tail -3
