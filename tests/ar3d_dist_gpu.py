"""Multi-GPU parity of the advection_reaction_3D kernels (run under torchrun on N GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29573 tests/ar3d_dist_gpu.py

Every rank builds the SAME global state from a seed, keeps its x-slab on its GPU, evaluates the
right-hand side (the upwind face travels over NVLink peer memory inside the kernel) many times
back to back -- no reduction in between, so the acknowledge protocol is what keeps the two
halo buffers consistent -- and checks its slab BIT-EXACT against the CPU oracle evaluated with
the neighbour's plane as halo (the reference's arithmetic for --npxyz N 1 1).
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "apps" / "advection_reaction_3D"))

import ar3d_oracle as orc  # noqa: E402
import run as ar  # noqa: E402


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64).ravel()


def main():
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    import datetime

    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lrank}"), timeout=datetime.timedelta(seconds=120))
    ctx = ar.make_context(lrank, rank, world)
    fails = 0
    cases = [(16, 0.01, 0, 0), (16, 0.01, 1, 0), (16, 0.01, 0, 1), (16, -0.3, 0, 0), (10, 0.5, 0, 0), (32, 0.01, 5, 0),
             (world, 0.01, 0, 0), (4 * world + 1, 0.01, 0, 0)]
    for n, c, chunk, generic in cases:
        p = orc.params(c=c)
        d = p["xmax"] / n
        i0, i1 = n * rank // world, n * (rank + 1) // world
        plan = ar.Plan(ctx, npts=n, c=c, planes_per_cta=chunk, force_generic=generic)
        assert plan.neq_loc == (i1 - i0) * n * n * 3
        ic = torch.empty(plan.neq_loc, dtype=torch.float64, device="cuda")
        plan.set_ic(ic)
        ok = np.array_equal(bits(ic.cpu().numpy()), bits(orc.initial_condition(n, p, rank, world)))
        f = torch.empty_like(ic)
        reps = 7
        for rep in range(reps):
            rng = np.random.default_rng(1000 + 17 * n + rep)
            y = rng.uniform(0.5, 1.5, (n, n, n, 3))
            up = (i0 - 1) % n if c > 0 else i1 % n
            want = orc.advection_reaction(y[i0:i1], p, d, d, d, halo=y[up])
            yd = torch.from_numpy(y[i0:i1].ravel()).cuda()
            which = ar.RHS_ADVECTION_REACTION if rep % 2 == 0 else ar.RHS_ADVECTION
            if which == ar.RHS_ADVECTION:
                want = orc.advection(y[i0:i1], c, d, d, d, halo=y[up])
            plan.rhs(which, yd, f)
            # no synchronisation with the other ranks here: the next call reuses the buffers
            ok = ok and np.array_equal(bits(f.cpu().numpy()), bits(want))
        plan.close()
        t = torch.tensor([0 if ok else 1], device="cuda")
        dist.all_reduce(t)
        if rank == 0:
            print(f"npts={n} c={c} chunk={chunk} generic={generic} fast={plan.fast}: {'ok' if t.item() == 0 else 'FAIL'}",
                  flush=True)
        fails += int(t.item())
    dist.barrier()
    if rank == 0:
        print("AR3D DIST OK" if fails == 0 else f"AR3D DIST FAILED ({fails})", flush=True)
    dist.destroy_process_group()
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
