"""Tuning sweep: achieved HBM GB/s of representative kernels vs launch geometry.
Run on the B200 box:  python tools/sweep.py [--n 24] > gpurun_out/sweep.txt
Timing: CUDA events on the launching stream, 3 warm-ups, operands rotated over
3 buffer sets (each op touches >= 256 MiB > 126 MB L2)."""
import argparse
import itertools
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from sundials_b200 import nvector as nv  # noqa: E402


def timed(fn, sets, reps):
    for s in sets:
        fn(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        fn(sets[r % len(sets)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=24)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--nv", type=int, default=8)
    a = ap.parse_args()
    n = 1 << a.n
    ctx = nv.default_context()
    nvv = a.nv
    sets = []
    for s in range(3):
        vs = [nv.N_VMake(torch.rand(n, dtype=torch.float64, device="cuda") + 0.5, ctx) for _ in range(2 * nvv + 2)]
        sets.append(vs)
    c = [0.3 + 0.1 * i for i in range(nvv)]
    ops = {
        "linear_sum": (24, lambda v: nv.N_VLinearSum(0.3, v[0], -2.1, v[1], v[2])),
        "axpy": (24, lambda v: nv.N_VLinearSum(0.3, v[0], 1.0, v[1], v[1])),
        "scale": (16, lambda v: nv.N_VScale(2.5, v[0], v[1])),
        "const": (8, lambda v: nv.N_VConst(1.5, v[0])),
        "div": (24, lambda v: nv.N_VDiv(v[0], v[1], v[2])),
        "dot_prod": (16, lambda v: nv.N_VDotProd(v[0], v[1])),
        "max_norm": (8, lambda v: nv.N_VMaxNorm(v[0])),
        "wrms_mask": (24, lambda v: nv.N_VWrmsNormMask(v[0], v[1], v[2])),
        "lin_comb": (8 * (nvv + 1), lambda v: nv.N_VLinearCombination(c, v[:nvv], v[nvv])),
        "scale_add_multi": (8 * (2 * nvv + 1), lambda v: nv.N_VScaleAddMulti(c, v[2 * nvv], v[:nvv], v[nvv:2 * nvv])),
        "dot_prod_multi": (8 * (nvv + 1), lambda v: nv.N_VDotProdMulti(v[nvv], v[:nvv])),
        "wrms_va": (16 * nvv, lambda v: nv.N_VWrmsNormVectorArray(v[:nvv], v[nvv:2 * nvv])),
        "linear_sum_va": (24 * nvv, lambda v: nv.N_VLinearSumVectorArray(0.3, v[:nvv], -2.1, v[nvv:2 * nvv], v[:nvv])),
    }
    print(f"# n=2^{a.n} nv={nvv} reps={a.reps}", flush=True)
    grid = list(itertools.product([4, 2], [4, 2, 1], [148 * 2, 148 * 4, 148 * 8, 148 * 16, 4096]))
    for name, (bpe, fn) in ops.items():
        best = None
        for (w, u, mb) in grid:
            ctx.set_tuning("vec_width", w)
            ctx.set_tuning("unroll", u)
            ctx.set_tuning("max_blocks", mb)
            t = timed(fn, sets, a.reps)
            gbs = bpe * n / t / 1e9
            print(json.dumps({"op": name, "W": w, "U": u, "max_blocks": mb, "us": round(t * 1e6, 2), "GBs": round(gbs, 1)}),
                  flush=True)
            if best is None or gbs > best[0]:
                best = (gbs, w, u, mb)
        print(f"# BEST {name}: {best[0]:.1f} GB/s W={best[1]} U={best[2]} max_blocks={best[3]}", flush=True)
    # torch copy for calibration (same definition as MEASURED_PEAKS: read+write bytes)
    a_, b_ = sets[0][0].data, sets[0][1].data
    t = timed(lambda s: b_.copy_(a_), sets, a.reps)
    print(f"# torch copy_: {16 * n / t / 1e9:.1f} GB/s", flush=True)


if __name__ == "__main__":
    main()
