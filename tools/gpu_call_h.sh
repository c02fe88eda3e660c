#!/bin/bash
# 2-GPU pass: advection_reaction_3D across ranks (kernel parity back to back, integrator runs), bench N=2
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ar3d_gpu.py -x -q -k "two_ranks or adams" > gpurun_out/pytest_ar3d_n2.log 2>&1; echo "pytest ar3d n2 rc=$?" | tee -a gpurun_out/pytest_ar3d_n2.log
tail -25 gpurun_out/pytest_ar3d_n2.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29573 tests/ar3d_dist_gpu.py > gpurun_out/ar3d_dist_n2.log 2>&1; echo "dist rc=$?"; grep -v "^W\|^\*\*\*" gpurun_out/ar3d_dist_n2.log | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
head -c 3000 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
