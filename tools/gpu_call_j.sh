#!/bin/bash
# N-GPU diagnostic (default 4) of the advection_reaction_3D strong-scaling run: where do the ranks wait?
# device-side counters (globaltimer) of the reduction exchange, the halo plane and the acknowledges, per rank,
# with both reduction transports; then the RHS kernels event-timed.
set -u
N=${1:-4}
mkdir -p gpurun_out
R=apps/advection_reaction_3D/run.py
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
A="--npts 512 --method ARK-IMEX --nls newton --fused --tf 0.05 --nout 1 --quiet --json"
timeout 150 $TR --master-port 29581 $R $A --profile > gpurun_out/ar3d_prof_n$N.json 2> gpurun_out/ar3d_prof_n$N.err; echo "profile p2p rc=$?"
grep PROFILE gpurun_out/ar3d_prof_n$N.err; tail -1 gpurun_out/ar3d_prof_n$N.json | head -c 600; echo
timeout 150 $TR --master-port 29582 $R $A --profile --p2p 0 > gpurun_out/ar3d_prof_nccl_n$N.json 2> gpurun_out/ar3d_prof_nccl_n$N.err; echo "profile nccl rc=$?"
grep PROFILE gpurun_out/ar3d_prof_nccl_n$N.err; tail -1 gpurun_out/ar3d_prof_nccl_n$N.json | head -c 600; echo
B200_AR3D_TIME_RHS=1 timeout 150 $TR --master-port 29583 $R $A > gpurun_out/ar3d_timerhs_n$N.json 2> gpurun_out/ar3d_timerhs_n$N.err; echo "time rhs rc=$?"
tail -1 gpurun_out/ar3d_timerhs_n$N.json | head -c 600; echo
