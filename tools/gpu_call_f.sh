#!/bin/bash
# 1-GPU pass: diffusion same-box comparison with the reference's CUDA benchmark, ncu of the TMA RHS kernel,
# cvDiurnal_kry timing, then the full round (tests, smoke, bench, reference arm, launch list)
set -u
mkdir -p gpurun_out
timeout 900 python tools/d2d_compare.py > gpurun_out/d2d_compare.json 2> gpurun_out/d2d_compare.err; echo "d2d_compare rc=$?"
cat gpurun_out/d2d_compare.json | head -c 1500; echo; tail -3 gpurun_out/d2d_compare.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_diffusion_rhs_tma -s 4 -c 1 -f -o gpurun_out/prof_rhs_tma python tools/rhs_bench.py --variants tma --rows 0 --forcing 1 --reps 3 > gpurun_out/prof_rhs_tma.log 2>&1; echo "ncu rc=$?"
( cd /tmp; for k in serial b200; do s=$(date +%s.%N); /root/repo/oracle/_ref/bin/cvDiurnal_kry_$k > /dev/null 2>&1; e=$(date +%s.%N); echo "cvDiurnal_kry_$k wall_s $(echo "$e - $s" | bc)"; done ) > gpurun_out/cvdiurnal_time.txt 2>&1; cat gpurun_out/cvdiurnal_time.txt
bash tools/gpu_round.sh
