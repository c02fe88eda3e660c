"""CVODE's fused-kernel plugin boundary on one B200: libsundials_cvode_fused_b200.so (one kernel per function)
beside (a) the unfused path on the same vector -- the reference's cvode_fused_stubs.c driving NVECTOR_B200 through
the ops table, i.e. what CVODE does with the fused kernels off -- and (b) the reference's own CUDA implementation
(cvode_fused_gpu.cpp, recompiled for sm_100a) on the reference's nvector_cuda.

Per function: `reps` calls back to back on vectors of n = 2^log2n doubles (128 MiB each at 24: every operand
streams from HBM), CUDA events around the batch on the legacy default stream (which both vectors use here).  GB/s = algorithmic bytes (each distinct operand read once, each
result written once) / time; frac = GB/s / the measured HBM peak (MEASURED_PEAKS.json, else the 6546.9 GB/s
fallback of the profiling guide).

    python tools/cvfused_bench.py [--log2n 24] [--reps 20] > gpurun_out/<tag>_cvfused_bench.json
"""
import argparse
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from sundials_b200.plugin import B200Plugin  # noqa: E402

STUBS = ROOT / "baseline" / "_ref" / "lib" / "libsundials_cvode_fused_stubs.so"
FUSED = ROOT / "sundials_b200" / "lib" / "libsundials_cvode_fused_b200.so"
_V, _D, _I = C.c_void_p, C.c_double, C.c_int
SIGS = {  # src/cvode/cvode_impl.h:639-672
    "cvEwtSetSS_fused": [_I, _D, _D, _V, _V, _V],
    "cvEwtSetSV_fused": [_I, _D, _V, _V, _V, _V],
    "cvCheckConstraints_fused": [_V, _V, _V, _V, _V],
    "cvNlsResid_fused": [_D, _D, _V, _V, _V, _V],
    "cvDiagSetup_formY": [_D, _D, _V, _V, _V, _V, _V],
    "cvDiagSetup_buildM": [_D, _D, _D, _V, _V, _V, _V, _V, _V, _V],
    "cvDiagSolve_updateM": [_D, _V],
}

def _bind(lib):
    for name, args in SIGS.items():
        f = getattr(lib, name)
        f.restype, f.argtypes = C.c_int, args
    return lib

UR = float(np.finfo(np.float64).eps)
# name -> (vectors, arrays moved per element by the fused kernel, arrays moved by the op sequence, call)
CASES = {
    "cvEwtSetSS_fused": (3, 3, 8, lambda L, v: L.cvEwtSetSS_fused(0, 1e-4, 1e-6, v[0], v[1], v[2])),
    "cvEwtSetSV_fused": (4, 4, 7, lambda L, v: L.cvEwtSetSV_fused(0, 1e-4, v[3], v[0], v[1], v[2])),
    "cvCheckConstraints_fused": (5, 5, 14, lambda L, v: L.cvCheckConstraints_fused(v[0], v[1], v[2], v[3], v[4])),
    "cvNlsResid_fused": (4, 4, 6, lambda L, v: L.cvNlsResid_fused(0.37, -0.013, v[0], v[1], v[2], v[3])),
    "cvDiagSetup_formY": (5, 5, 6, lambda L, v: L.cvDiagSetup_formY(0.02, 0.05, v[0], v[1], v[2], v[3], v[4])),
    "cvDiagSetup_buildM": (7, 8, 28, lambda L, v: L.cvDiagSetup_buildM(0.1, UR, 0.02, v[0], v[1], v[2], v[3], v[4], v[5], v[6])),
    "cvDiagSolve_updateM": (1, 2, 8, lambda L, v: L.cvDiagSolve_updateM(1.0, v[0])),  # r = 1: M keeps its magnitude
}


def timed(fn, reps, stream):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


def run(log2n=24, reps=20, with_ref_cuda=True):
    """with_ref_cuda=False touches nothing under oracle/ (bench.py's leg): the unfused arm is the reference's
    stubs library from baseline/_ref driving NVECTOR_B200"""
    a = argparse.Namespace(log2n=log2n, reps=reps, no_refcuda=not with_ref_cuda)
    n = 1 << a.log2n
    pk = ROOT / "MEASURED_PEAKS.json"
    peaks = json.loads(pk.read_text()) if pk.exists() else {}
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"

    core = bench.load_reference() if with_ref_cuda else bench.load_host()  # the host framework (RTLD_GLOBAL)
    stubs, fused = _bind(C.CDLL(str(STUBS))), _bind(C.CDLL(str(FUSED)))
    P = B200Plugin()
    rng = np.random.default_rng(7)
    init = [rng.uniform(0.5, 2.0, n) for _ in range(2)]

    def b200_vec(k):
        h = P.new(n, None, P.DEVICE, fused=True)
        P.host(h, n)[...] = init[k % 2]
        P.to_device(h)
        P.drop_host(h)
        return h

    vb = [b200_vec(k) for k in range(7)]
    sb = torch.cuda.default_stream()  # the default context runs on the legacy default stream, like nvector_cuda

    out = {"n": n, "reps": a.reps, "hbm_peak_GBs": peak, "hbm_peak_source": peak_src, "functions": {}}
    for name, (nv, fa, ua, call) in CASES.items():
        t_f = timed(lambda: call(fused, vb), a.reps, sb)
        t_u = timed(lambda: call(stubs, vb), a.reps, sb)
        gb = fa * 8 * n / 1e9
        out["functions"][name] = {
            "fused_us": round(t_f, 2), "fused_GBs": round(gb / (t_f * 1e-6), 1), "fused_frac_of_peak": round(gb / (t_f * 1e-6) / peak, 3),
            "fused_bytes_per_elt": 8 * fa, "unfused_us": round(t_u, 2), "unfused_bytes_per_elt": 8 * ua,
            "speedup_vs_unfused": round(t_u / t_f, 2),
        }
    for h in vb:
        P.Destroy(h)

    if not a.no_refcuda:
        so = ROOT / "oracle" / "_ref" / "lib"
        cu = C.CDLL(str(so / "libsundials_nveccuda_ref.so"), mode=C.RTLD_GLOBAL)
        refk = _bind(C.CDLL(str(so / "libsundials_cvode_fused_cuda_ref.so")))
        cu.N_VNew_Cuda.restype, cu.N_VNew_Cuda.argtypes = C.c_void_p, [C.c_int64, C.c_void_p]
        cu.N_VCopyToDevice_Cuda.argtypes = [C.c_void_p]
        core.N_VDestroy.argtypes = [C.c_void_p]
        sctx = C.c_void_p()
        assert core.SUNContext_Create(0, C.byref(sctx)) == 0

        def cuda_vec(k):
            v = cu.N_VNew_Cuda(n, sctx)
            assert v, "N_VNew_Cuda failed"
            np.ctypeslib.as_array(core.N_VGetArrayPointer(v), shape=(n,))[...] = init[k % 2]
            cu.N_VCopyToDevice_Cuda(v)
            return v

        vc = [cuda_vec(k) for k in range(7)]
        s0 = torch.cuda.default_stream()
        for name, (nv, fa, ua, call) in CASES.items():
            t_r = timed(lambda: call(refk, vc), a.reps, s0)
            f = out["functions"][name]
            f["ref_cuda_fused_us"] = round(t_r, 2)
            f["speedup_vs_ref_cuda_fused"] = round(t_r / f["fused_us"], 2)
        for v in vc:
            core.N_VDestroy(v)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=24)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--no-refcuda", action="store_true")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    torch.cuda.init()
    print(json.dumps(run(a.log2n, a.reps, not a.no_refcuda), indent=1))


if __name__ == "__main__":
    main()
