#!/bin/bash
# 2-GPU pass: diffusion_2D two-rank parity (halo rows over peer memory), weak-scaling timing, bench at N=2
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_diffusion2d_gpu.py -x -q -k "two_ranks" > gpurun_out/pytest_diffusion_n2.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_diffusion_n2.log
R=apps/diffusion_2D/run.py
timeout 200 python $R --nx 8192 --ny 8192 --tf 1e-4 --nout 1 --output 0 --json > gpurun_out/d2d_w_n1.json 2> gpurun_out/d2d_w_n1.err; echo "N=1 rc=$?"; tail -1 gpurun_out/d2d_w_n1.json | head -c 700; echo
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 $R --nx 8192 --ny 16384 --yu 2 --tf 1e-4 --nout 1 --output 0 --json > gpurun_out/d2d_w_n2.json 2> gpurun_out/d2d_w_n2.err; echo "N=2 rc=$?"; tail -1 gpurun_out/d2d_w_n2.json | head -c 700; echo; tail -3 gpurun_out/d2d_w_n2.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], json.dumps(d.get('diffusion_2D'))[:900])
PY
tail -3 gpurun_out/bench_n2.err
