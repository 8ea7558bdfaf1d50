#!/bin/bash
# 8-GPU: the C4 bench (8 M atoms) -- bench.py itself runs the decomposed-vs-single check before its timed region.
# Everything under a tight timeout.  BOX8=stacked adds the run with the eight cubes stacked along the slab axis.
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 300 --warmup 100 --no-cpu 2>gpurun_out/b8.err > gpurun_out/b8.json; python scripts/summ.py "N=8 dd" < gpurun_out/b8.json
grep -v "^W0\|OMP_NUM\|^\*\*\*\|^$" gpurun_out/b8.err | tail -4
if [ "${BOX8:-}" = stacked ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 8 --steps 300 --warmup 100 --no-cpu --no-e2e --box stacked 2>gpurun_out/b8s.err > gpurun_out/b8s.json; python scripts/summ.py "N=8 dd stacked" < gpurun_out/b8s.json
fi
