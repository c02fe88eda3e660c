"""BASELINE configs[2]: the N_Vector length sweep, 2^16 ... 2^30 doubles per GPU, one
representative op per kernel class, NVECTOR_B200 vs the reference's nvector_serial /
nvector_openmp on the box's host cores (bounded lengths).

    python tools/length_sweep.py [--max 30] > gpurun_out/length_sweep.json

Timing: CUDA events on the context's stream around `reps` back-to-back calls after 3
warm-ups; operands rotate over `sets` disjoint buffer sets.  "l2_resident": the bytes one
op touches x sets fit the 126 MB L2, so the figure is an L2 / launch-latency number, not
an HBM number (every length <= 2^20 and some 2^22 cases).  Reductions return their scalar
to the host (the N_Vector API), so their time includes the host hand-off.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from sundials_b200 import nvector as nv  # noqa: E402

L2_BYTES = 126e6


def ops_for(nvv):
    c = [0.3 + 0.1 * i for i in range(nvv)]
    return {
        "N_VLinearSum": (24, 3, lambda v: nv.N_VLinearSum(0.3, v[0], -2.1, v[1], v[2])),
        "N_VScale": (16, 2, lambda v: nv.N_VScale(2.5, v[0], v[1])),
        "N_VConst": (8, 1, lambda v: nv.N_VConst(1.5, v[0])),
        "N_VDotProd": (16, 2, lambda v: nv.N_VDotProd(v[0], v[1])),
        "N_VMaxNorm": (8, 1, lambda v: nv.N_VMaxNorm(v[0])),
        "N_VWrmsNormMask": (24, 3, lambda v: nv.N_VWrmsNormMask(v[0], v[1], v[2])),
        f"N_VLinearCombination(nv={nvv})": (8 * (nvv + 1), nvv + 1,
                                            lambda v: nv.N_VLinearCombination(c, v[:nvv], v[nvv])),
        f"N_VScaleAddMulti(nv={nvv})": (8 * (2 * nvv + 1), 2 * nvv + 1,
                                        lambda v: nv.N_VScaleAddMulti(c, v[2 * nvv], v[:nvv], v[nvv:2 * nvv])),
        f"N_VDotProdMulti(nv={nvv})": (8 * (nvv + 1), nvv + 1, lambda v: nv.N_VDotProdMulti(v[nvv], v[:nvv])),
    }


def cpu_sweep(lengths, threads):
    """nvector_serial (1 core) and nvector_openmp (all cores) of the unmodified reference"""
    import ctypes as C

    import bench
    from sundials_b200.plugin import Api

    lib = bench.load_reference()
    sctx = C.c_void_p()
    assert lib.SUNContext_Create(0, C.byref(sctx)) == 0
    api = Api(lib, "")
    out = {}
    for L in lengths:
        n = 1 << L
        for kind in ("serial", "openmp"):
            def new():
                if kind == "serial":
                    v = lib.N_VNew_Serial(n, sctx)
                    lib.N_VEnableFusedOps_Serial(v, 1)
                else:
                    v = lib.N_VNew_OpenMP(n, threads, sctx)
                    lib.N_VEnableFusedOps_OpenMP(v, 1)
                np.ctypeslib.as_array(lib.N_VGetArrayPointer(v), shape=(n,))[...] = 0.75
                return v
            vs = [new() for _ in range(10)]
            X = api.varray(vs[:8])
            c = api.coefs([0.3 + 0.1 * i for i in range(8)])
            cases = {"N_VLinearSum": (24, lambda: api.LinearSum(0.3, vs[0], -2.1, vs[1], vs[2])),
                     "N_VDotProd": (16, lambda: api.DotProd(vs[0], vs[1])),
                     "N_VLinearCombination(nv=8)": (72, lambda: api.LinearCombination(8, c, X, vs[8]))}
            for name, (bpe, fn) in cases.items():
                fn()
                reps = max(2, min(200, int(2e9 / (bpe * n))))
                t0 = time.perf_counter()
                for _ in range(reps):
                    fn()
                dt = (time.perf_counter() - t0) / reps
                out.setdefault(f"2^{L}", {}).setdefault(name, {})[kind] = {"us": round(dt * 1e6, 1),
                                                                          "GBs": round(bpe * n / dt / 1e9, 2)}
            lib.N_VDestroy.argtypes = [C.c_void_p]
            for v in vs:
                lib.N_VDestroy(v)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min", type=int, default=16)
    ap.add_argument("--max", type=int, default=30)
    ap.add_argument("--step", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    ctx = nv.default_context()
    peak = 6650.0
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peak = float(json.loads(pk.read_text()).get("hbm_gbs", peak))
    free, total = torch.cuda.mem_get_info()
    res = {"gpu": torch.cuda.get_device_name(0), "peak_GBs": peak, "lengths": {}}
    for L in range(a.min, a.max + 1, a.step):
        n = 1 << L
        nvv = 8 if L <= 29 else 4
        need = 2 * nvv + 1
        budget = int(min(free * 0.8, 100e9))
        sets = max(1, min(8, budget // (need * n * 8)))
        if need * n * 8 > budget:
            res["lengths"][f"2^{L}"] = {"skipped": f"{need} vectors of {n * 8 / 2**30:.0f} GiB exceed the budget"}
            continue
        pool = [[nv.N_VMake(torch.full((n,), 0.75 + 0.01 * i, dtype=torch.float64, device="cuda"), ctx)
                 for i in range(need)] for _ in range(sets)]
        row = {}
        for name, (bpe, nops, fn) in ops_for(nvv).items():
            for s in pool[:3]:
                fn(s)
            torch.cuda.synchronize()
            reps = int(max(5, min(200, 20e9 / (bpe * n))))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for r in range(reps):
                fn(pool[r % sets])
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / reps * 1e3
            gbs = bpe * n / us / 1e3
            row[name] = {"us": round(us, 2), "GBs": round(gbs, 1), "frac_of_peak": round(gbs / peak, 3),
                         "l2_resident": bool(bpe * n * sets < L2_BYTES)}
        res["lengths"][f"2^{L}"] = {"nvecs": nvv, "sets": int(sets), "ops": row}
        del pool
        torch.cuda.empty_cache()
        print(f"# 2^{L} done", file=sys.stderr, flush=True)
    if not a.no_cpu:
        import os

        threads = os.cpu_count() or 1
        res["cpu"] = {"cores": threads, "kind": "reference nvector_serial (1 core) / nvector_openmp (all cores)",
                      "lengths": cpu_sweep([L for L in range(a.min, min(a.max, 24) + 1, 4)], threads)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
