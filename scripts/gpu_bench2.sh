#!/bin/bash
# 2-GPU: decomposition parity check (uneven layer split), then the decomposed bench.  Everything under a tight timeout.
DD_NCELL=30 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dd_check.py 2>&1 | grep "dd_check\|differ\|rror" | tail -3
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 100 --no-cpu 2>gpurun_out/b2.err > gpurun_out/b2.json; python scripts/summ.py "N=2 dd" < gpurun_out/b2.json
grep -v "^W0\|OMP_NUM\|^\*\*\*\|^$" gpurun_out/b2.err | tail -3
