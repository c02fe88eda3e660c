#!/usr/bin/env python
"""Kernel timings of apps/advection_reaction_3D on one GPU (CUDA events on the context's stream).

    python tools/ar3d_bench.py [--npts 320] [--reps 20] [--chunks 4,8,16,32] [--json out.json]

Operands rotate over more vectors than the 126 MB L2 holds.  Algorithmic bytes: RHS 16 per
unknown (y read once, ydot written once), block solve 24 per unknown (y, b read, x written).
"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "apps" / "advection_reaction_3D"))
import run as ar  # noqa: E402


def time_op(fn, reps, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--npts", type=int, default=320)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--chunks", default="0")
    ap.add_argument("--generic", action="store_true")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    ctx = ar.make_context(0, 0, 1)
    # the vector library's kernels run on the legacy default stream unless told otherwise, and so
    # do torch's events on its current (default) stream
    n = a.npts
    neq = 3 * n ** 3
    nbuf = max(3, int(1.5e9 // (neq * 8)) + 1)
    ys = [1.0 + torch.rand(neq, dtype=torch.float64, device="cuda") for _ in range(nbuf)]
    fs = [torch.empty(neq, dtype=torch.float64, device="cuda") for _ in range(nbuf)]
    res = {"npts": n, "unknowns": neq, "buffers": nbuf, "rows": []}
    for chunk in [int(c) for c in a.chunks.split(",")]:
        plan = ar.Plan(ctx, npts=n, planes_per_cta=chunk, force_generic=1 if a.generic else 0)
        rows = []
        for name, which in (("advection", ar.RHS_ADVECTION), ("reaction", ar.RHS_REACTION),
                            ("advection_reaction", ar.RHS_ADVECTION_REACTION)):
            t = time_op(lambda i: plan.rhs(which, ys[i % nbuf], fs[i % nbuf]), a.reps)
            rows.append((name, t, 16.0 * neq / t / 1e9))
        t = time_op(lambda i: plan.psolve(ys[i % nbuf], fs[i % nbuf], fs[(i + 1) % nbuf], 1e-3), a.reps)
        rows.append(("block_solve", t, 24.0 * neq / t / 1e9))
        for name, t, gbs in rows:
            print(f"npts={n} fast={plan.fast} planes_per_cta={chunk or 'default'} {name:20s} {t * 1e6:10.1f} us {gbs:8.1f} GB/s",
                  flush=True)
            res["rows"].append({"kernel": name, "fast": plan.fast, "planes_per_cta": chunk, "us": t * 1e6, "GBps": gbs})
        plan.close()
    if a.json:
        Path(a.json).write_text(json.dumps(res, indent=1) + "\n")


if __name__ == "__main__":
    main()
