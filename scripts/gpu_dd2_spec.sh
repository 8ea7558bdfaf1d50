#!/bin/bash
# 2 GPUs: the launch sent ahead in decomposed runs (spec_force=2, experimental) -- one short run under a tight timeout;
# the error text names the wait that gave up
mkdir -p gpurun_out
SEPGPU_OPTS="spec_force=2" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 100 --warmup 20 --equilibrate 60 --no-cpu --no-e2e 2>gpurun_out/b2s.err > gpurun_out/b2s.json; echo "rc=$?"; python scripts/summ.py "N=2 dd spec" < gpurun_out/b2s.json; grep -a "never arrived" gpurun_out/b2s.err | head -2 | cut -c1-300
