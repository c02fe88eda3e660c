"""Where a scalar-returning N_Vector op's time goes, on an otherwise idle GPU (fresh process, nothing
else running): API time (C driver loop of apps/nvector_perf, CUDA events), device time (%globaltimer
from the first CTA to the publication, "profile" stamps) and their difference (launch path + PCIe
hand-off + host poll), per length.  Companion of tools/mb_reduce2.cu (the same measurement on
stand-alone kernel variants).

    python tools/reduce_latency.py [--log2n 16 18 20 22 24 26] > gpurun_out/reduce_latency.json
"""
import argparse
import ctypes as C
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from sundials_b200 import _lib  # noqa: E402
from sundials_b200.plugin import B200Plugin  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, nargs="+", default=[16, 18, 20, 22, 24, 26])
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--pdl", type=int, default=1)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    lib, P, perf = _lib.load(), B200Plugin(), bench.load_perf()
    ctx = C.c_void_p()
    _lib.check(lib.b200vec_ctx_create(C.byref(ctx), 0, None), "ctx_create")
    lib.b200vec_ctx_set_tuning(ctx, b"pdl", a.pdl)
    ops = ["N_VDotProd", "N_VMaxNorm", "N_VMin", "N_VL1Norm", "N_VWrmsNorm", "N_VWrmsNormMask", "N_VInvTest",
           "N_VConstrMask", "N_VMinQuotient", "N_VDotProdMulti", "N_VWrmsNormVectorArray"]
    out = {"gpu": torch.cuda.get_device_name(0), "pdl": a.pdl, "lengths": {}}
    for L in a.log2n:
        n = 1 << L
        cnt = [0]

        def newvec():
            cnt[0] += 1
            v = P.new(n, ctx, P.DEVICE, fused=True)
            P.Const(0.75 + 0.01 * (cnt[0] % 17), v)
            return v

        X, Y, Z = [newvec() for _ in range(8)], [newvec() for _ in range(8)], [newvec() for _ in range(8)]
        vec = {"X": X, "Y": Y, "Z": Z, "S": newvec(), "T": newvec(), "W": newvec(), "ID": newvec(), "CN": newvec(),
               "YY": [Y], "ZZ": [Z]}
        suite = bench.Suite(perf, vec)
        row = {}
        for name in ops:
            i = suite.names.index(name)
            suite.op(i, 3)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            suite.op(i, a.reps)
            e1.record()
            torch.cuda.synchronize()
            api = e0.elapsed_time(e1) / a.reps * 1e3
            lib.b200vec_ctx_set_tuning(ctx, b"profile", 1)
            devs = []
            for _ in range(5):
                lib.b200vec_ctx_set_tuning(ctx, b"prof_stamp_reset", 1)
                suite.op(i, 1)
                t0 = lib.b200vec_ctx_get_tuning(ctx, b"prof_counter_6")
                t1 = lib.b200vec_ctx_get_tuning(ctx, b"prof_counter_7")
                if 0 < t0 < t1:
                    devs.append((t1 - t0) * 1e-3)
            lib.b200vec_ctx_set_tuning(ctx, b"profile", 0)
            dev = sorted(devs)[len(devs) // 2] if devs else None
            bpe = suite.bpe[i]
            row[name] = {"api_us": round(api, 2), "dev_us": round(dev, 2) if dev else None,
                         "launch_and_return_us": round(api - dev, 2) if dev else None,
                         "api_GBs": round(bpe * n / api / 1e3, 1), "dev_GBs": round(bpe * n / dev / 1e3, 1) if dev else None}
        out["lengths"][f"2^{L}"] = row
        for v in X + Y + Z + [vec["S"], vec["T"], vec["W"], vec["ID"], vec["CN"]]:
            P.Destroy(v)
        lib.b200vec_ctx_sync(ctx)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
