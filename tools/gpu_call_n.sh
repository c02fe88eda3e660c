#!/bin/bash
# 1-GPU pass: parity of the bucketed multi-output reduction kernels, Gram-Schmidt timing, reduction per-op timings
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -x -q > gpurun_out/pytest_parity.log 2>&1; echo "pytest parity rc=$?"; tail -3 gpurun_out/pytest_parity.log
timeout 300 python -m pytest tests/test_reference_harness_gpu.py -x -q -k "gram or harness_tiny or 1000" > gpurun_out/pytest_harness.log 2>&1; echo "pytest harness rc=$?"; tail -3 gpurun_out/pytest_harness.log
timeout 200 python tools/gs_bench.py --maxl 5 --reps 5 > gpurun_out/gs_bench_maxl5_v2.json 2> gpurun_out/gs_bench_maxl5_v2.err; echo "gs5 rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/gs_bench_maxl5_v2.json'))
for a in ('b200','ref_cuda'):
    for g in ('classical','modified'):
        print(a,g,d[a][g]['cycle_us'],d[a][g]['per_k_us'])
P
timeout 300 python bench.py --steps 5 --no-cpu-baseline --no-diffusion --no-ar3d > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_short.json'))
print(d['value'], d['ms_per_step'])
for k in ('N_VDotProdMulti','N_VWrmsNormVectorArray','N_VWrmsNormMaskVectorArray','N_VDotProdMultiLocal'):
    print(k, d['per_op'][k])
print(d['gram_schmidt']['b200']['classical']['cycle_us'], d['gram_schmidt']['b200']['modified']['cycle_us'])
P
