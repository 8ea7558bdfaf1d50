#!/bin/bash
# 2-GPU check of the domain decomposition (run with: gpurun --gpus 2 -- bash scripts/gpu_dd2.sh); uneven layer split
DD_NCELL=30 timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dd_check.py 2>&1 | grep "dd_check\|differ\|rror" | tail -3
