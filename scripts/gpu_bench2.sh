#!/bin/bash
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dd_check.py 2>&1 | grep "dd_check\|differ\|rror\|PASS\|FAIL" | tail -5
for cfg in "1 0"; do
set -- $cfg
SEPGPU_DD_P2P=$1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 100 --overlap $2 2>gpurun_out/b2.err > gpurun_out/b2_$1$2.json; python scripts/summ.py "N=2 dd p2p=$1 overlap=$2" < gpurun_out/b2_$1$2.json
done
grep -v "^W0\|OMP_NUM\|^\*\*\*\|^$" gpurun_out/b2.err | tail -3
