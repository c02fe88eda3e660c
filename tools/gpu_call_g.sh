#!/bin/bash
# 1-GPU pass for apps/advection_reaction_3D: parity tests, kernel timings, one integrator run
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ar3d_gpu.py -x -q > gpurun_out/pytest_ar3d.log 2>&1; echo "pytest ar3d rc=$?" | tee -a gpurun_out/pytest_ar3d.log
tail -25 gpurun_out/pytest_ar3d.log
timeout 300 python tools/ar3d_bench.py --npts 320 --chunks 4,8,16,32 --json gpurun_out/ar3d_bench_320.json > gpurun_out/ar3d_bench.log 2>&1; echo "ar3d_bench rc=$?"
timeout 120 python tools/ar3d_bench.py --npts 320 --generic >> gpurun_out/ar3d_bench.log 2>&1
cat gpurun_out/ar3d_bench.log
timeout 300 python apps/advection_reaction_3D/run.py --npts 128 --method ARK-IMEX --nls newton --tf 0.1 --nout 2 --fused --json > gpurun_out/ar3d_run_128.log 2>&1; echo "run rc=$?"; tail -14 gpurun_out/ar3d_run_128.log
