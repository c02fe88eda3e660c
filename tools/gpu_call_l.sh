#!/bin/bash
# final 1-GPU pass: bench line of the final code, ncu --set full of the advection_reaction_3D kernels
set -u
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; head -c 300 gpurun_out/bench.json; echo; tail -2 gpurun_out/bench.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_ar3d_march|k_ar3d_reaction|k_ar3d_psolve' -c 16 -f -o gpurun_out/prof_ar3d \
  python tools/ar3d_bench.py --npts 320 --reps 1 > gpurun_out/prof_ar3d.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/prof_ar3d.ncu-rep --page raw --csv > gpurun_out/prof_ar3d_raw.csv 2>/dev/null; echo "export rc=$?"; wc -c gpurun_out/prof_ar3d_raw.csv
rm -f gpurun_out/prof_ar3d.ncu-rep.tmp
