#!/bin/bash
# 1-GPU pass: ncu --set full of the diffusion RHS kernel, length sweep 2^16..2^30
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_diffusion_rhs -s 4 -c 1 -f -o gpurun_out/prof_rhs python tools/rhs_bench.py --variants la4x2 --rows 0 --forcing 1 --reps 3 > gpurun_out/prof_rhs.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/prof_rhs.log
timeout 1200 python tools/length_sweep.py > gpurun_out/length_sweep.json 2> gpurun_out/length_sweep.err; echo "sweep rc=$?"
tail -5 gpurun_out/length_sweep.err; head -c 400 gpurun_out/length_sweep.json
