#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_parity_gpu.py -x -q > gpurun_out/pytest_parity.log 2>&1; echo "pytest parity rc=$?"; tail -3 gpurun_out/pytest_parity.log
timeout 100 python tools/gs_ops_breakdown.py > gpurun_out/gs_ops_v2.json 2> gpurun_out/gs_ops_v2.err; echo rc=$?; grep "^{" gpurun_out/gs_ops_v2.err
