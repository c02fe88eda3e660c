#!/bin/bash
# 1-GPU pass: new parity tests (2^30 indexing), smoke, full bench (suite + diffusion_2D + advection_reaction_3D
# + Gram-Schmidt legs), reference arm, Gram-Schmidt vs the reference's nvector_cuda, advection_reaction_3D
# size / tf sweep (separates one-time from per-step cost), ncu launch list of one advection_reaction_3D run.
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_parity_gpu.py -x -q > gpurun_out/pytest_parity.log 2>&1; echo "pytest parity rc=$?"; tail -3 gpurun_out/pytest_parity.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; head -c 600 gpurun_out/bench.json; echo; tail -3 gpurun_out/bench.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; head -c 400 gpurun_out/bench_ref.json; echo
timeout 200 python tools/gs_bench.py --maxl 5 --reps 5 > gpurun_out/gs_bench_maxl5.json 2> gpurun_out/gs_bench_maxl5.err; echo "gs5 rc=$?"; head -c 1200 gpurun_out/gs_bench_maxl5.json; echo
timeout 200 python tools/gs_bench.py --maxl 20 --reps 3 --cpu-log2n 20 > gpurun_out/gs_bench_maxl20.json 2> gpurun_out/gs_bench_maxl20.err; echo "gs20 rc=$?"
R=apps/advection_reaction_3D/run.py
for cfg in "128 0.05" "256 0.05" "256 0.1" "384 0.05"; do
  set -- $cfg
  timeout 120 python $R --npts $1 --method ARK-IMEX --nls newton --fused --tf $2 --nout 1 --quiet --json > gpurun_out/ar3d_n1_npts$1_tf$2.json 2> gpurun_out/ar3d_n1_npts$1_tf$2.err; echo "ar3d npts=$1 tf=$2 rc=$?"
  tail -1 gpurun_out/ar3d_n1_npts$1_tf$2.json | head -c 700; echo
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/ar3d_launches.csv python $R --npts 256 --method ARK-IMEX --nls newton --fused --tf 0.004 --nout 1 --quiet > gpurun_out/ar3d_under_ncu.log 2>&1; echo "ncu rc=$?"
