#!/bin/bash
# 1-GPU pass: A/B of the reduction kernels' L2 prefetch look-ahead, then the whole GPU test-suite.
set -u
mkdir -p gpurun_out
timeout 300 python tools/reduce_ab.py --log2n 22,24,26 --reps 40 > gpurun_out/reduce_ab.json 2> gpurun_out/reduce_ab.err; echo "reduce_ab rc=$?"
grep "log2n': 24" gpurun_out/reduce_ab.err
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_gpu.log
