#!/bin/bash
# 2-GPU check of the domain decomposition (run with: gpurun --gpus 2 -- bash scripts/gpu_dd2.sh)
nvidia-smi -L | head -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dd_check.py 2>&1 | grep -v "^W0\|OMP_NUM\|^\*\*\*" | tail -25
echo "exit: ${PIPESTATUS[0]}"
