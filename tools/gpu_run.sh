#!/bin/bash
# tools/gpu_run.sh <tag> <what...> -- ONE parameterised script for the round's gpurun calls
# (replaces round 1's seventeen one-shot gpu_call_*.sh).  Runs on the GPU box from the repo root;
# everything lands in gpurun_out/<tag>_*.  what:
#   tests [pytest args]   pytest -m gpu (default: the whole suite)
#   bench [bench args]    python bench.py ...            -> <tag>_bench.json
#   mb    [log2n ...]     build/mb_reduce2               -> <tag>_mb_reduce2.txt
#   launches [bench args] ncu launch list of a short bench run (gpu__time_duration)
#   full  <script args>   ncu --set full of tools/profile_kernels.py
#   cmd   <shell command> anything else
set -u
tag=$1; shift
what=$1; shift
mkdir -p gpurun_out
case "$what" in
  tests)   timeout 1500 python -m pytest tests -m gpu -x -q "$@" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${tag}_pytest.log ;;
  bench)   timeout 1500 python bench.py "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -c 1800 gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err ;;
  mb)      timeout 600 build/mb_reduce2 "$@" > gpurun_out/${tag}_mb_reduce2.txt 2>&1; echo "mb rc=$?"; tail -3 gpurun_out/${tag}_mb_reduce2.txt ;;
  launches) timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py "$@" > gpurun_out/${tag}_launches_bench.log 2>&1; echo "ncu rc=$?"; wc -l gpurun_out/${tag}_launches.csv ;;
  full)    timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_map|k_reduce|k_lincomb|k_scaleadd' -s 14 -c 14 -o gpurun_out/${tag}_full -f python tools/profile_kernels.py "$@" > gpurun_out/${tag}_full.log 2>&1; echo "ncu rc=$?"; ls -la gpurun_out/${tag}_full.ncu-rep ;;
  cmd)     bash -c "$*" ;;
  *) echo "unknown: $what"; exit 2 ;;
esac
