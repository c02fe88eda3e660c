#!/bin/bash
# multi-GPU pass (N = $1 ranks): distributed parity + bench at N
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tests/dist_parity_gpu.py > gpurun_out/dist_parity_$N.log 2>&1; echo "dist parity rc=$?"
tail -4 gpurun_out/dist_parity_$N.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
head -c 700 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
