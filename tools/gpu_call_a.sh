#!/bin/bash
# 1-GPU pass: diffusion parity (regression of the RHS kernel), RHS variants, reference nvector_cuda head-to-head
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_diffusion2d_gpu.py -x -q > gpurun_out/pytest_diffusion.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_diffusion.log
timeout 600 python tools/rhs_bench.py > gpurun_out/rhs_bench.txt 2>&1; echo "rhs_bench rc=$?"
cat gpurun_out/rhs_bench.txt | tail -45
timeout 900 python tools/ref_cuda_suite.py > gpurun_out/ref_cuda_suite.json 2> gpurun_out/ref_cuda_suite.err; echo "ref_cuda rc=$?"
tail -3 gpurun_out/ref_cuda_suite.err; head -c 600 gpurun_out/ref_cuda_suite.json
