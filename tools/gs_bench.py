"""Gram-Schmidt orthogonalisation on the new vector: the reference's
UNMODIFIED SUNClassicalGS / SUNModifiedGS (src/sundials/sundials_iterative.c:45-170, from the
host framework baseline/_ref/lib/libsundials_host.so), plus the fused SUNClassicalGS_B200, building a Krylov basis column by column exactly as SPGMR does
(sunlinsol_spgmr.c: orthogonalise v[k] against v[0..k-1], normalise by the returned norm),
timed per call on

  * NVECTOR_B200 (fused ops enabled),
  * the reference's own nvector_cuda recompiled for sm_100a (same box, fused ops enabled), if
    oracle/_ref/lib/libsundials_nveccuda_ref.so exists,
  * nvector_openmp on the host cores (bounded length).

Every GS call ends in a scalar-returning N_VDotProd, so host wall clock around the call is
the device time plus the launch/round-trip path an integrator experiences.

Ideal HBM traffic per call at column k (each distinct operand once per op):
  classical: N_VDotProdMulti(k+1) 8N(k+1) + N_VLinearCombination(k+1, in place) 8N(k+2)
             + N_VDotProd(v_k, v_k) 8N                                   = 8N(2k+4)
  modified : N_VDotProd(v_k,v_k) 8N + k x (N_VDotProd 16N + N_VLinearSum 24N) + 8N = 8N(5k+2)

    python tools/gs_bench.py [--log2n 24] [--maxl 5] [--reps 5] > gpurun_out/gs_bench.json
"""
import argparse
import ctypes as C
import json
import os
import statistics
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

V = C.c_void_p
dp = C.POINTER(C.c_double)


def ideal_bytes(gs, k, n):
    if gs == "fused_classical":   # multi-dot 8N(k+1) + in-place combination with the norm in its epilogue 8N(k+2)
        return 8 * n * (2 * k + 3)
    if gs == "fused_modified":    # 2-wide multi-dot 16N + (k-1) x axpy+dot 32N + last update with the norm 24N
        return 8 * n * (4 * k + 1)
    return 8 * n * (2 * k + 4) if gs == "classical" else 8 * n * (5 * k + 2)


def bind_core(core):
    core.SUNClassicalGS.restype = C.c_int
    core.SUNClassicalGS.argtypes = [C.POINTER(V), C.POINTER(dp), C.c_int, C.c_int, dp, dp, C.POINTER(V)]
    core.SUNModifiedGS.restype = C.c_int
    core.SUNModifiedGS.argtypes = [C.POINTER(V), C.POINTER(dp), C.c_int, C.c_int, dp]
    core.N_VScale.restype, core.N_VScale.argtypes = None, [C.c_double, V, V]
    core.N_VDotProd.restype, core.N_VDotProd.argtypes = C.c_double, [V, V]
    core.N_VDestroy.restype, core.N_VDestroy.argtypes = None, [V]


def time_gs(core, newvec, fill, sync, n, maxl, reps, fused=None):
    """Returns {gs: {"per_k_us": [...], "cycle_us": total of k = 1..maxl, "h_last": ...}}.
    S[k]: fixed pseudo-random source columns; V[k]: the basis being built."""
    S = [newvec() for _ in range(maxl + 1)]
    Vv = [newvec() for _ in range(maxl + 1)]
    for i, s in enumerate(S):
        fill(s, 100 + i)
    basis = (V * (maxl + 1))(*Vv)
    vtemp = (V * (maxl + 1))()
    stemp = (C.c_double * (maxl + 1))()
    rows = [(C.c_double * maxl)() for _ in range(maxl + 1)]
    h = (dp * (maxl + 1))(*[C.cast(r, dp) for r in rows])
    nrm = C.c_double()
    out = {}
    for gs in ("classical", "modified") + (("fused_classical", "fused_modified") if fused is not None else ()):
        per_k = {k: [] for k in range(1, maxl + 1)}
        for rep in range(reps + 1):                      # first pass = warm-up
            core.N_VScale(1.0, S[0], Vv[0])
            d = core.N_VDotProd(Vv[0], Vv[0])
            core.N_VScale(1.0 / d ** 0.5, Vv[0], Vv[0])
            for k in range(1, maxl + 1):
                core.N_VScale(1.0, S[k], Vv[k])
                sync()
                t0 = time.perf_counter()
                if gs == "classical":
                    rc = core.SUNClassicalGS(basis, h, k, maxl, C.byref(nrm), stemp, vtemp)
                elif gs == "fused_classical":
                    rc = fused[0](basis, h, k, maxl, C.byref(nrm), stemp, vtemp)
                elif gs == "fused_modified":
                    rc = fused[1](basis, h, k, maxl, C.byref(nrm))
                else:
                    rc = core.SUNModifiedGS(basis, h, k, maxl, C.byref(nrm))
                dt = time.perf_counter() - t0
                assert rc == 0 and nrm.value > 0.0, (gs, k, rc, nrm.value)
                core.N_VScale(1.0 / nrm.value, Vv[k], Vv[k])
                if rep > 0:
                    per_k[k].append(dt * 1e6)
        med = [statistics.median(per_k[k]) for k in range(1, maxl + 1)]
        # orthogonality of the finished basis: |<v_i, v_maxl>| for i < maxl
        orth = max(abs(core.N_VDotProd(Vv[i], Vv[maxl])) for i in range(maxl))
        out[gs] = {
            "per_k_us": [round(u, 2) for u in med],
            "per_k_GBs": [round(ideal_bytes(gs, k, n) / med[k - 1] / 1e3, 1) for k in range(1, maxl + 1)],
            "cycle_us": round(sum(med), 2),
            "cycle_GBs": round(sum(ideal_bytes(gs, k, n) for k in range(1, maxl + 1)) / sum(med) / 1e3, 1),
            "cycle_GBs_at_unfused_traffic": round(sum(ideal_bytes(gs.replace("fused_", ""), k, n)
                                                      for k in range(1, maxl + 1)) / sum(med) / 1e3, 1),
            "h_last_column": [rows[i][maxl - 1] for i in range(maxl)],
            "last_norm": nrm.value,
            "max_abs_dot_with_last": orth,
        }
    for v in S + Vv:
        core.N_VDestroy(v)
    return out


def lcg_fill(a, seed):
    """values in [-1, 1] from a cheap seeded generator (numpy, vectorised)"""
    import numpy as np

    a[...] = np.random.default_rng(seed).uniform(-1.0, 1.0, a.shape[0])


def run(log2n=24, maxl=5, reps=5, cpu_log2n=22, with_ref_cuda=True, with_cpu=True, b200_ctx=None, rank=0, world=1):
    """world > 1 (bench.py under torchrun; b200_ctx carries the communicator): the basis vectors
    are distributed (contiguous block of n per rank), every reduction inside GS folds the ranks'
    partials -- all ranks call this collectively; the single-GPU baselines are skipped."""
    import numpy as np
    import torch

    import bench
    from sundials_b200 import _lib
    from sundials_b200.plugin import B200Plugin

    n = 1 << log2n
    # NVECTOR_B200 needs only the unmodified host framework (baseline/_ref); the reference's CPU / CUDA
    # vectors (oracle/_ref) are loaded further down, for the baselines alone
    core = bench.load_host()
    bind_core(core)
    sctx = C.c_void_p()
    assert core.SUNContext_Create(0, C.byref(sctx)) == 0
    res = {"log2n": log2n, "maxl": maxl, "reps": reps, "n_gpus": world,
           "timing": "host wall clock around each SUN*GS call (device idle before; the call ends in a "
                     "scalar-returning N_VDotProd), median over reps",
           "ideal_bytes": "classical 8N(2k+4), modified 8N(5k+2) per call at column k"}

    # ---- NVECTOR_B200
    P = B200Plugin()
    lib = _lib.load()
    ctx = b200_ctx
    if ctx is None:
        ctx = C.c_void_p()
        _lib.check(lib.b200vec_ctx_create(C.byref(ctx), torch.cuda.current_device(), None), "ctx_create")

    def new_b200():
        v = lib.N_VNewWithCtx_B200(n, P.DEVICE, ctx, sctx)
        assert v, "N_VNewWithCtx_B200 failed"
        lib.N_VEnableFusedOps_B200(v, 1)
        if world > 1:
            assert lib.N_VMakeDistributed_B200(v, n * world) == 0
        return v

    def fill_b200(v, seed):
        lcg_fill(P.host(v, n), seed + 1000 * rank)
        P.to_device(v)
        P.drop_host(v)

    lib.SUNClassicalGS_B200.restype = C.c_int
    lib.SUNClassicalGS_B200.argtypes = [C.POINTER(V), C.POINTER(dp), C.c_int, C.c_int, dp, dp, C.POINTER(V)]
    lib.SUNModifiedGS_B200.restype = C.c_int
    lib.SUNModifiedGS_B200.argtypes = [C.POINTER(V), C.POINTER(dp), C.c_int, C.c_int, dp]
    res["b200"] = time_gs(core, new_b200, fill_b200, torch.cuda.synchronize, n, maxl, reps,
                          fused=(lib.SUNClassicalGS_B200, lib.SUNModifiedGS_B200))
    res["ideal_bytes"] += ("; fused_classical (SUNClassicalGS_B200: 2 kernels per column) 8N(2k+3); fused_modified "
                           "(SUNModifiedGS_B200: k + 1 kernels per column) 8N(4k+1)")

    # ---- reference nvector_cuda on the same GPU
    so = ROOT / "oracle" / "_ref" / "lib" / "libsundials_nveccuda_ref.so"
    if with_ref_cuda and world == 1 and so.exists():
        core = bench.load_reference()
        bind_core(core)
        cu = C.CDLL(str(so), mode=C.RTLD_GLOBAL)
        cu.N_VNew_Cuda.restype, cu.N_VNew_Cuda.argtypes = V, [C.c_int64, V]
        cu.N_VEnableFusedOps_Cuda.restype, cu.N_VEnableFusedOps_Cuda.argtypes = C.c_int, [V, C.c_int]
        cu.N_VCopyToDevice_Cuda.argtypes = [V]

        def new_cuda():
            v = cu.N_VNew_Cuda(n, sctx)
            assert v, "N_VNew_Cuda failed"
            cu.N_VEnableFusedOps_Cuda(v, 1)
            return v

        def fill_cuda(v, seed):
            lcg_fill(np.ctypeslib.as_array(core.N_VGetArrayPointer(v), shape=(n,)), seed)
            cu.N_VCopyToDevice_Cuda(v)

        res["ref_cuda"] = time_gs(core, new_cuda, fill_cuda, torch.cuda.synchronize, n, maxl, reps)
        # same inputs, same routine: the Hessenberg columns agree to rounding
        for gs in ("classical", "modified"):
            a, b = res["b200"][gs]["h_last_column"], res["ref_cuda"][gs]["h_last_column"]
            res["b200"][gs]["h_max_abs_diff_vs_ref_cuda"] = max(abs(x - y) for x, y in zip(a, b))
            res[f"speedup_{gs}_cycle"] = round(res["ref_cuda"][gs]["cycle_us"] / res["b200"][gs]["cycle_us"], 2)

    # ---- nvector_openmp on the host cores, bounded length
    if with_cpu and world == 1:
        core = bench.load_reference()
        bind_core(core)
        threads = os.cpu_count() or 1
        nc = 1 << min(cpu_log2n, log2n)

        def new_omp():
            v = core.N_VNew_OpenMP(nc, threads, sctx)
            core.N_VEnableFusedOps_OpenMP(v, 1)
            return v

        def fill_omp(v, seed):
            lcg_fill(np.ctypeslib.as_array(core.N_VGetArrayPointer(v), shape=(nc,)), seed)

        r = time_gs(core, new_omp, fill_omp, lambda: None, nc, maxl, max(2, reps // 2))
        r["cores"], r["log2n"] = threads, min(cpu_log2n, log2n)
        res["openmp_cpu"] = r
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=24)
    ap.add_argument("--maxl", type=int, default=5, help="Krylov dimension (SPGMR default 5)")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cpu-log2n", type=int, default=22)
    ap.add_argument("--no-baselines", action="store_true", help="NVECTOR_B200 only (e.g. under ncu)")
    a = ap.parse_args()
    import torch

    torch.cuda.set_device(0)
    torch.cuda.init()
    out = run(a.log2n, a.maxl, a.reps, a.cpu_log2n, with_ref_cuda=not a.no_baselines, with_cpu=not a.no_baselines)
    out["gpu"] = torch.cuda.get_device_name(0)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
