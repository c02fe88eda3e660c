#!/bin/bash
# 2-GPU pass: bench.py at N = 2 on the final code (distributed Gram-Schmidt leg, per-leg clocks)
set -u
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
head -c 400 gpurun_out/bench_n2.json; echo; grep -v "^W\|^\*\*\*" gpurun_out/bench_n2.err | tail -5
