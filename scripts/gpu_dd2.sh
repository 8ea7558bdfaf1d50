#!/bin/bash
# 2-GPU check of the domain decomposition (run with: gpurun --gpus 2 -- bash scripts/gpu_dd2.sh); even and odd layer counts
for nc in 28 30; do
DD_NCELL=$nc timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dd_check.py 2>&1 | grep "dd_check\|differ\|rror" | tail -3
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 100 2>gpurun_out/b2.err > gpurun_out/b2.json; python scripts/summ.py "N=2 dd" < gpurun_out/b2.json
grep -v "^W0\|OMP_NUM\|^\*\*\*\|^$" gpurun_out/b2.err | tail -3
