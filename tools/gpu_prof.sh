#!/bin/bash
# ncu --set full capture of the representative kernels (one GPU) + reduction microbench
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'k_map|k_reduce|k_lincomb|k_scaleadd' -s 9 -c 9 -f -o gpurun_out/prof_r01 \
  python tools/profile_kernels.py --n 24 --reps 2 > gpurun_out/prof_r01.log 2>&1; echo "ncu rc=$?"
timeout 600 build/mb_reduce 24 > gpurun_out/mb_reduce_24.txt 2>&1; echo "mb rc=$?"
tail -3 gpurun_out/prof_r01.log; head -8 gpurun_out/mb_reduce_24.txt
