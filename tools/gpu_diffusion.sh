#!/bin/bash
# GPU-box pass for the re-hosted diffusion_2D benchmark: parity tests vs the reference's
# goldens, then bounded 8192^2 runs (timing), RHS-kernel timing, and an ncu launch list.
set -u
N=${1:-1}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_diffusion2d_gpu.py -x -q > gpurun_out/pytest_diffusion.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_diffusion.log
R=apps/diffusion_2D/run.py
for tf in 1e-6 1e-5 1e-4; do
  timeout 200 python $R --nx 8192 --ny 8192 --tf $tf --nout 1 --output 0 --json > gpurun_out/d2d_8192_tf$tf.json 2> gpurun_out/d2d_8192_tf$tf.err; echo "8192^2 tf=$tf rc=$?"
  tail -1 gpurun_out/d2d_8192_tf$tf.json | head -c 900; echo
done
B200_DIFFUSION_TIME_RHS=1 timeout 200 python $R --nx 8192 --ny 8192 --tf 1e-5 --nout 1 --output 0 --json > gpurun_out/d2d_8192_rhs.json 2>&1; echo "rhs timing rc=$?"
tail -1 gpurun_out/d2d_8192_rhs.json | head -c 900; echo
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/d2d_launches.csv python $R --nx 8192 --ny 8192 --tf 1e-6 --nout 1 --output 0 > gpurun_out/d2d_under_ncu.log 2>&1; echo "ncu rc=$?"
if [ "$N" -gt 1 ]; then
  for tf in 1e-5 1e-4; do
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 $R --nx 8192 --ny $((8192*N)) --yu $N --tf $tf --nout 1 --output 0 --json > gpurun_out/d2d_8192_n${N}_tf$tf.json 2> gpurun_out/d2d_8192_n${N}_tf$tf.err; echo "N=$N tf=$tf rc=$?"
    tail -1 gpurun_out/d2d_8192_n${N}_tf$tf.json | head -c 900; echo
  done
fi
