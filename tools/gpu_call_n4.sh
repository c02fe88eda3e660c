#!/bin/bash
# N-GPU pass (default 4): distributed-vector parity on N ranks (both transports), advection_reaction_3D
# kernel parity with distinct west/east neighbours, diffusion_2D with inner strips (two halo rows per
# launch), then bench.py at N.  Charged N x box time: kept short.
set -u
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tests/dist_parity_gpu.py > gpurun_out/dist_parity_n$N.log 2>&1; echo "dist parity rc=$?"
grep -v "^W\|^\*\*\*" gpurun_out/dist_parity_n$N.log | tail -4
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29573 tests/ar3d_dist_gpu.py > gpurun_out/ar3d_dist_n$N.log 2>&1; echo "ar3d dist rc=$?"
grep -v "^W\|^\*\*\*" gpurun_out/ar3d_dist_n$N.log | tail -10
timeout 200 python -m pytest tests/test_diffusion2d_gpu.py -x -q -k "n_ranks and $N" > gpurun_out/pytest_diffusion_n$N.log 2>&1; echo "pytest diffusion rc=$?"
tail -5 gpurun_out/pytest_diffusion_n$N.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"
head -c 2500 gpurun_out/bench_n$N.json; grep -v "^W\|^\*\*\*" gpurun_out/bench_n$N.err | tail -5
