#!/bin/bash
# 2 GPUs: the launch sent ahead in decomposed runs (spec_force=2, experimental) -- short, under a tight timeout
mkdir -p gpurun_out
for eq in 60 300; do
SEPGPU_OPTS="spec_force=2" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 100 --warmup 20 --equilibrate $eq --no-cpu --no-e2e 2>gpurun_out/b2s.err > gpurun_out/b2s.json; echo "equilibrate=$eq rc=$?"; python scripts/summ.py "N=2 dd spec" < gpurun_out/b2s.json; grep -a "RuntimeError\|never arrived" gpurun_out/b2s.err | head -2
done
