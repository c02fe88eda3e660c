"""Same-box baseline: the reference's OWN CUDA vector (nvector_cuda, unmodified, recompiled
for sm_100a by oracle/Makefile `refcuda`) against NVECTOR_B200 on the bench suite.

Both arms run the identical 55-op suite (bench.make_suite: benchmarks/nvector cases, fused
ops ENABLED on both), through the generic N_V* dispatch of the reference core for
nvector_cuda and the N_V*_B200 ops-table functions for ours; same seeded inputs, same
timing: per op 1 warm-up + `reps` calls bracketed by device synchronisation, host wall
clock (what an integrator experiences: launch path + kernel + any host round trip).

    python tools/ref_cuda_suite.py [--log2n 24] [--reps 10] > gpurun_out/ref_cuda_suite.json
"""
import argparse
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from sundials_b200.plugin import Api, B200Plugin  # noqa: E402


def time_suite(suite, reps):
    out = {}
    for name, bpe, fn in suite:
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        out[name] = (time.perf_counter() - t0) / reps * 1e6
    # whole step, back to back
    for _ in range(2):
        for _, _, fn in suite:
            fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for _, _, fn in suite:
            fn()
    torch.cuda.synchronize()
    return out, (time.perf_counter() - t0) / reps * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=24)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    n = 1 << a.log2n
    torch.cuda.set_device(0)
    torch.cuda.init()

    # ---- arm 1: reference nvector_cuda
    core = bench.load_reference()
    so = ROOT / "oracle" / "_ref" / "lib" / "libsundials_nveccuda_ref.so"
    cu = C.CDLL(str(so), mode=C.RTLD_GLOBAL)
    cu.N_VNew_Cuda.restype, cu.N_VNew_Cuda.argtypes = C.c_void_p, [C.c_int64, C.c_void_p]
    cu.N_VEnableFusedOps_Cuda.restype, cu.N_VEnableFusedOps_Cuda.argtypes = C.c_int, [C.c_void_p, C.c_int]
    cu.N_VCopyToDevice_Cuda.argtypes = [C.c_void_p]
    core.N_VDestroy.argtypes = [C.c_void_p]
    sctx = C.c_void_p()
    assert core.SUNContext_Create(0, C.byref(sctx)) == 0

    def new_cuda():
        v = cu.N_VNew_Cuda(n, sctx)
        assert v, "N_VNew_Cuda failed"
        assert cu.N_VEnableFusedOps_Cuda(v, 1) == 0
        return v

    vec = bench.alloc_vectors(new_cuda)
    rng = np.random.default_rng(1234)
    bench._init_values(vec, lambda v: np.ctypeslib.as_array(core.N_VGetArrayPointer(v), shape=(n,)), rng, n)
    for v in bench.all_handles(vec):
        cu.N_VCopyToDevice_Cuda(v)
    suite, res_ref, _k1 = bench.make_suite(Api(core, ""), vec)
    ref_us, ref_ms = time_suite(suite, a.reps)
    bpes = {name: bpe for name, bpe, _ in suite}
    bytes_per_step = sum(bpes.values()) * n
    res_ref = dict(res_ref)
    for v in bench.all_handles(vec):
        core.N_VDestroy(v)
    del suite, vec
    torch.cuda.synchronize()

    # ---- arm 2: NVECTOR_B200
    P = B200Plugin()
    from sundials_b200 import _lib

    ctx = C.c_void_p()
    _lib.check(_lib.load().b200vec_ctx_create(C.byref(ctx), 0, None), "ctx_create")
    vec = bench.alloc_vectors(lambda: P.new(n, ctx, P.DEVICE, True))
    rng = np.random.default_rng(1234)
    # same values as arm 1: fill in the same order from the same seed, one vector at a time
    # (the pinned host mirror is dropped after the upload)
    hs = bench.all_handles(vec)
    for v in hs:
        P.host(v, n)[...] = rng.uniform(0.5, 1.5, n) * (rng.integers(0, 2, n) * 2 - 1)
    P.host(vec["W"], n)[...] = rng.uniform(0.5, 1.5, n)
    P.host(vec["ID"], n)[...] = rng.integers(0, 2, n).astype(np.float64)
    P.host(vec["CN"], n)[...] = rng.integers(-2, 3, n).astype(np.float64)
    for v in hs:
        P.to_device(v)
        P.drop_host(v)
    suite, res_b200, _k2 = bench.make_suite(P, vec)
    b_us, b_ms = time_suite(suite, a.reps)

    rows = {}
    for name in bpes:
        rows[name] = {"B_per_elt": bpes[name], "ref_cuda_us": round(ref_us[name], 2), "b200_us": round(b_us[name], 2),
                      "ref_cuda_GBs": round(bpes[name] * n / ref_us[name] / 1e3, 1),
                      "b200_GBs": round(bpes[name] * n / b_us[name] / 1e3, 1),
                      "speedup": round(ref_us[name] / b_us[name], 2)}
    # scalar results of the two arms (same inputs): reductions agree to rounding
    agree = {k: [res_ref.get(k), res_b200.get(k)] for k in res_b200}
    print(json.dumps({"log2n": a.log2n, "reps": a.reps, "gpu": torch.cuda.get_device_name(0),
                      "timing": "host wall clock around reps calls, device synchronised on both sides",
                      "suite_ms": {"ref_cuda": round(ref_ms, 3), "b200": round(b_ms, 3)},
                      "suite_GBs": {"ref_cuda": round(bytes_per_step / ref_ms / 1e6, 1),
                                    "b200": round(bytes_per_step / b_ms / 1e6, 1)},
                      "suite_speedup": round(ref_ms / b_ms, 2), "per_op": rows, "scalars_ref_vs_b200": agree}))


if __name__ == "__main__":
    main()
