#!/bin/bash
# 2 GPUs: decomposition tests (C ABI under torchrun incl. butane, SEP_NGPU through the sep_* API), then the decomposed benches
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dd.py -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/dd2.txt
if [ "${DD2_BENCH:-1}" = 1 ]; then
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 100 --no-cpu 2>gpurun_out/b2.err > gpurun_out/b2.json; python scripts/summ.py "N=2 dd" < gpurun_out/b2.json
grep -v "^W0\|OMP_NUM\|^\*\*\*\|^$" gpurun_out/b2.err | tail -3
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --workload butane --gpus 2 --steps 200 --warmup 40 2>gpurun_out/b2_butane.err > gpurun_out/b2_butane.json; tail -c 1500 gpurun_out/b2_butane.json; tail -3 gpurun_out/b2_butane.err
fi
