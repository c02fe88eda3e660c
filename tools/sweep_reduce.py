"""Kernel-only sweep of the reduction kernels (result_host = NULL: no host sync
inside the timed region).  python tools/sweep_reduce.py --n 24"""
import argparse
import ctypes as C
import itertools
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from sundials_b200 import nvector as nv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=24)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--quick", action="store_true", help="dot/max-norm only: grid cap x PDL on/off, kernel-only and API")
    a = ap.parse_args()
    n = 1 << a.n
    ctx = nv.default_context()
    L = ctx.lib
    nvv = 8
    sets = [[torch.rand(n, dtype=torch.float64, device="cuda") + 0.5 for _ in range(2 * nvv + 1)] for _ in range(3)]

    def tab(ts):
        return (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])

    ops = {
        "dot_prod": (16, lambda v: L.b200vec_dot_prod(ctx.h, v[0].data_ptr(), v[1].data_ptr(), n, None)),
        "max_norm": (8, lambda v: L.b200vec_max_norm(ctx.h, v[0].data_ptr(), n, None)),
        "wsqr_mask": (24, lambda v: L.b200vec_wsqr_sum_mask(ctx.h, v[0].data_ptr(), v[1].data_ptr(), v[2].data_ptr(), n, None)),
        "inv_test": (16, lambda v: L.b200vec_inv_test(ctx.h, v[0].data_ptr(), v[1].data_ptr(), n, None)),
        "dot_prod_multi8": (8 * (nvv + 1), lambda v: L.b200vec_dot_prod_multi(ctx.h, nvv, v[2 * nvv].data_ptr(), tab(v[:nvv]), n, None)),
        "wsqr_va8": (16 * nvv, lambda v: L.b200vec_wsqr_sum_vector_array(ctx.h, nvv, tab(v[:nvv]), tab(v[nvv:2 * nvv]), None, n, None)),
        "wsqr_mask_va8": (8 * (2 * nvv + 1), lambda v: L.b200vec_wsqr_sum_vector_array(ctx.h, nvv, tab(v[:nvv]), tab(v[nvv:2 * nvv]), v[2 * nvv].data_ptr(), n, None)),
    }
    if a.quick:
        out = C.c_double()
        api = {
            "dot_prod": (16, lambda v: L.b200vec_dot_prod(ctx.h, v[0].data_ptr(), v[1].data_ptr(), n, C.byref(out))),
            "max_norm": (8, lambda v: L.b200vec_max_norm(ctx.h, v[0].data_ptr(), n, C.byref(out))),
        }
        for pdl in (1, 0):
            ctx.set_tuning("pdl", pdl)
            for mb in (148, 296, 444, 592, 888, 1184):
                ctx.set_tuning("max_blocks", mb)
                for mode, table in (("kernel", ops), ("api", api)):
                    for name in ("dot_prod", "max_norm"):
                        bpe, fn = table[name]
                        for s in sets:
                            fn(s)
                        torch.cuda.synchronize()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        for r in range(a.reps):
                            fn(sets[r % 3])
                        e1.record()
                        torch.cuda.synchronize()
                        t = e0.elapsed_time(e1) / a.reps * 1e-3
                        print(json.dumps({"op": name, "mode": mode, "pdl": pdl, "max_blocks": mb,
                                          "us": round(t * 1e6, 2), "GBs": round(bpe * n / t / 1e9, 1)}), flush=True)
        return
    grid = list(itertools.product([4, 2], [4, 2, 1], [148, 296, 592, 1184, 2368, 4096]))
    for name, (bpe, fn) in ops.items():
        best = None
        for (w, u, mb) in grid:
            ctx.set_tuning("vec_width", w)
            ctx.set_tuning("unroll", u)
            ctx.set_tuning("max_blocks", mb)
            for s in sets:
                fn(s)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for r in range(a.reps):
                fn(sets[r % 3])
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1) / a.reps * 1e-3
            gbs = bpe * n / t / 1e9
            print(json.dumps({"op": name, "W": w, "U": u, "max_blocks": mb, "us": round(t * 1e6, 2), "GBs": round(gbs, 1)}), flush=True)
            if best is None or gbs > best[0]:
                best = (gbs, w, u, mb)
        print(f"# BEST {name}: {best[0]:.1f} GB/s W={best[1]} U={best[2]} max_blocks={best[3]}", flush=True)


if __name__ == "__main__":
    main()
