#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/rhs_bench.py --variants la4x2,tma --rows 0,64,128 --forcing 1,0 > gpurun_out/rhs_bench_tma.txt 2>&1; echo "rhs_bench rc=$?"
cat gpurun_out/rhs_bench_tma.txt | tail -20
timeout 300 python tools/rhs_bench.py --nx 1000 --ny 777 --variants la4x2,tma --rows 0,5 --forcing 1 > gpurun_out/rhs_bench_tma_small.txt 2>&1; echo "rhs_bench small rc=$?"
cat gpurun_out/rhs_bench_tma_small.txt | tail -8
timeout 900 python -m pytest tests/test_diffusion2d_gpu.py -x -q > gpurun_out/pytest_diffusion.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_diffusion.log
