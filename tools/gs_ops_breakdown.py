"""Per-op wall time of one classical Gram-Schmidt column on NVECTOR_B200 (the three vector ops
SUNClassicalGS issues, sundials_iterative.c:133-150), device idle before each op, through the
ops-table functions: where does a CGS call's time go?

    python tools/gs_ops_breakdown.py [--log2n 24] [--maxl 5] [--reps 7] > gpurun_out/gs_ops.json
"""
import argparse
import ctypes as C
import json
import statistics
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from sundials_b200 import _lib  # noqa: E402
from sundials_b200.plugin import B200Plugin  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=24)
    ap.add_argument("--maxl", type=int, default=5)
    ap.add_argument("--reps", type=int, default=7)
    a = ap.parse_args()
    n = 1 << a.log2n
    torch.cuda.set_device(0)
    P = B200Plugin()
    lib = _lib.load()
    ctx = C.c_void_p()
    _lib.check(lib.b200vec_ctx_create(C.byref(ctx), 0, None), "ctx_create")
    Vv = []
    for i in range(a.maxl + 1):
        v = P.new(n, ctx, P.DEVICE, fused=True)
        P.host(v, n)[...] = np.random.default_rng(100 + i).uniform(-1, 1, n)
        P.to_device(v)
        P.drop_host(v)
        Vv.append(v)
    src = P.Clone(Vv[0])
    rows = []

    def timed(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e6

    for k in range(1, a.maxl + 1):
        nv = k + 1
        Y = P.varray(Vv[:k] + [Vv[k]])
        dots = (C.c_double * nv)()
        X = P.varray([Vv[k]] + Vv[:k])
        t = {"dot_prod_multi": [], "linear_combination": [], "dot_prod": [], "scale_copy": []}
        for rep in range(a.reps + 1):
            P.Scale(1.0, Vv[k], src)          # keep a copy: every repetition orthogonalises the same column
            r = [timed(lambda: P.DotProdMulti(nv, Vv[k], Y, dots))]
            c = P.coefs([1.0] + [-1e-3 * dots[i] for i in range(k)])   # small correction: the column stays well scaled
            r.append(timed(lambda: P.LinearCombination(nv, c, X, Vv[k])))
            r.append(timed(lambda: P.DotProd(Vv[k], Vv[k])))
            r.append(timed(lambda: P.Scale(1.0, src, Vv[k])))
            if rep > 0:
                for key, val in zip(t, r):
                    t[key].append(val)
        row = {"k": k, "nvec": nv}
        for key, bpe in (("dot_prod_multi", 8 * nv), ("linear_combination", 8 * (nv + 1)), ("dot_prod", 8), ("scale_copy", 16)):
            us = statistics.median(t[key])
            row[key + "_us"] = round(us, 1)
            row[key + "_GBs"] = round(bpe * n / us / 1e3, 0)
        row["sum_us"] = round(row["dot_prod_multi_us"] + row["linear_combination_us"] + row["dot_prod_us"], 1)
        rows.append(row)
        print(row, file=sys.stderr, flush=True)
    print(json.dumps({"log2n": a.log2n, "timing": "host wall clock, device synchronised before and after each op",
                      "bytes": "distinct operands once: multi-dot 8N*nvec, linear combination 8N*(nvec+1), dot(v,v) 8N",
                      "rows": rows}))


if __name__ == "__main__":
    main()
