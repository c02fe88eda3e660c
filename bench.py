#!/usr/bin/env python
"""bench.py -- N_Vector op throughput on B200 (BASELINE.json metric).

A *step* is one pass of the reference's N_Vector performance suite
(benchmarks/nvector/test_nvector_performance.c: every standard, reduction, fused
and vector-array op, 16 N_VLinearSum cases, 4 N_VScale cases, ...) over vectors
of length 2^LOG2N per GPU with nvecs=8, nsums=4, fused ops ENABLED, called
through the plugin boundary -- the `N_V*_B200` functions that sit in the
SUNDIALS N_Vector_Ops table (include/nvector_b200.h).  Throughput is
ALGORITHMIC bytes (SURVEY.md section 8d byte model: each distinct operand read
once, each output written once) per second, summed over the suite.

  value   suite GB/s with all operands resident in HBM (CUDA events, max over ranks)
  e2e     same suite, but every step first copies the nvecs input vectors from
          pinned host memory (N_VCopyToDevice_B200) and ends with a D2H read of a
          result vector plus the reduction scalars
  roofline  the dominant kernel by share of the step (k_scaleadd_rows<4>) timed
          with CUDA events, vs the measured HBM copy peak
  cpu_baseline  the reference's own nvector_openmp (all host threads) and
          nvector_serial (1 core) on a bounded sample of the same suite
  --impl reference : the reference CPU implementation alone (driver's baseline arm)

Multi-GPU (torchrun, one rank per GPU): weak scaling, each rank owns the local
block of a contiguous 1-D partition (MPIPlusX pattern); streaming ops need no
communication, every reduction does an NCCL allreduce of its scalars.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

NVECS, NSUMS = 8, 4


# --------------------------------------------------------------------------
# the suite: (name, algorithmic bytes per element, callable) -- identical for
# the B200 plugin (suffix _B200) and the reference generic dispatch (suffix "")
# --------------------------------------------------------------------------
def make_suite(api, vec):
    """vec: dict of handles: X[8], Y[8], Z[8], S, T, W, ID, CN, YY[4][8], ZZ[4][8]"""
    X, Y, Z, S, T, W, ID, CN = vec["X"], vec["Y"], vec["Z"], vec["S"], vec["T"], vec["W"], vec["ID"], vec["CN"]
    YY, ZZ = vec["YY"], vec["ZZ"]
    nv, ns = len(X), len(YY)
    a, b = 0.37, -1.63
    c8 = api.coefs([0.11 * (j + 1) * (-1) ** j for j in range(nv)])
    c8_one = api.coefs([1.0] + [0.11 * (j + 1) for j in range(1, nv)])
    c4 = api.coefs([0.21 * (j + 1) * (-1) ** j for j in range(ns)])
    cs8 = api.coefs([1.0 + 0.01 * j for j in range(nv)])
    aX, aY, aZ = api.varray(X), api.varray(Y), api.varray(Z)
    aYS = api.varray([S] + Y[1:])  # fused in-place forms act on the scratch vector
    aW = api.varray(Y)  # distinct weight vectors (byte model: 16 nv)
    aYY, aZZ = api.varray2d(YY), api.varray2d(ZZ)
    dots = (C.c_double * nv)()
    nrm = (C.c_double * nv)()
    res = {}
    s = []

    def op(name, bpe, fn):
        s.append((name, bpe, fn))

    # N_VLinearSum cases 1a..9 (test_nvector_performance.c:70-472); in-place forms
    # run on scratch copies so the read-only inputs X, Y survive the step
    op("N_VScale-2(copy)", 16, lambda: api.Scale(1.0, Y[0], S))
    op("N_VLinearSum-1a", 24, lambda: api.LinearSum(1.0, X[0], 1.0, S, S))
    op("N_VLinearSum-1b", 24, lambda: api.LinearSum(-1.0, X[0], 1.0, S, S))
    op("N_VLinearSum-1c", 24, lambda: api.LinearSum(a, X[0], 1.0, S, S))
    op("N_VLinearSum-2a", 24, lambda: api.LinearSum(1.0, S, 1.0, Y[0], S))
    op("N_VLinearSum-2b", 24, lambda: api.LinearSum(1.0, S, -1.0, Y[0], S))
    op("N_VLinearSum-2c", 24, lambda: api.LinearSum(1.0, S, b, Y[0], S))
    op("N_VLinearSum-3", 24, lambda: api.LinearSum(1.0, X[0], 1.0, Y[0], Z[0]))
    op("N_VLinearSum-4a", 24, lambda: api.LinearSum(1.0, X[1], -1.0, Y[1], Z[1]))
    op("N_VLinearSum-4b", 24, lambda: api.LinearSum(-1.0, X[2], 1.0, Y[2], Z[2]))
    op("N_VLinearSum-5a", 24, lambda: api.LinearSum(1.0, X[3], b, Y[3], Z[3]))
    op("N_VLinearSum-5b", 24, lambda: api.LinearSum(a, X[4], 1.0, Y[4], Z[4]))
    op("N_VLinearSum-6a", 24, lambda: api.LinearSum(-1.0, X[5], b, Y[5], Z[5]))
    op("N_VLinearSum-6b", 24, lambda: api.LinearSum(a, X[6], -1.0, Y[6], Z[6]))
    op("N_VLinearSum-7", 24, lambda: api.LinearSum(a, X[7], a, Y[7], Z[7]))
    op("N_VLinearSum-8", 24, lambda: api.LinearSum(a, X[0], -a, Y[1], Z[0]))
    op("N_VLinearSum-9", 24, lambda: api.LinearSum(a, X[1], b, Y[2], Z[1]))
    op("N_VConst", 8, lambda: api.Const(1.5, T))
    op("N_VProd", 24, lambda: api.Prod(X[2], Y[3], Z[2]))
    op("N_VDiv", 24, lambda: api.Div(X[3], Y[4], Z[3]))
    op("N_VScale-1(inplace)", 16, lambda: api.Scale(1.0009765625, S, S))
    op("N_VScale-3(neg)", 16, lambda: api.Scale(-1.0, X[4], Z[4]))
    op("N_VScale-4", 16, lambda: api.Scale(a, X[5], Z[5]))
    op("N_VAbs", 16, lambda: api.Abs(X[6], Z[6]))
    op("N_VInv", 16, lambda: api.Inv(X[7], Z[7]))
    op("N_VAddConst", 16, lambda: api.AddConst(X[0], b, Z[0]))
    op("N_VDotProd", 16, lambda: res.__setitem__("dot", api.DotProd(X[1], Y[1])))
    op("N_VMaxNorm", 8, lambda: res.__setitem__("max", api.MaxNorm(X[2])))
    op("N_VWrmsNorm", 16, lambda: res.__setitem__("wrms", api.WrmsNorm(X[3], W)))
    op("N_VWrmsNormMask", 24, lambda: res.__setitem__("wrmsmask", api.WrmsNormMask(X[4], W, ID)))
    op("N_VMin", 8, lambda: res.__setitem__("min", api.Min(X[5])))
    op("N_VWL2Norm", 16, lambda: res.__setitem__("wl2", api.WL2Norm(X[6], W)))
    op("N_VL1Norm", 8, lambda: res.__setitem__("l1", api.L1Norm(X[7])))
    op("N_VCompare", 16, lambda: api.Compare(0.75, X[0], Z[0]))
    op("N_VInvTest", 16, lambda: res.__setitem__("invtest", api.InvTest(X[1], Z[1])))
    op("N_VConstrMask", 24, lambda: res.__setitem__("constr", api.ConstrMask(CN, X[2], Z[2])))
    op("N_VMinQuotient", 16, lambda: res.__setitem__("minq", api.MinQuotient(X[3], Y[3])))
    # fused (test_nvector_performance.c:1312-1700); -1/-2 are the in-place forms
    op("N_VLinearCombination-1", 8 * (nv + 1), lambda: api.LinearCombination(nv, c8_one, aYS, S))
    op("N_VLinearCombination-2", 8 * (nv + 1), lambda: api.LinearCombination(nv, c8, aYS, S))
    op("N_VLinearCombination-3", 8 * (nv + 1), lambda: api.LinearCombination(nv, c8, aX, T))
    op("N_VScaleAddMulti-1", 8 * (2 * nv + 1), lambda: api.ScaleAddMulti(nv, c8, X[0], aZ, aZ))
    op("N_VScaleAddMulti-2", 8 * (2 * nv + 1), lambda: api.ScaleAddMulti(nv, c8, X[1], aY, aZ))
    op("N_VDotProdMulti", 8 * (nv + 1), lambda: api.DotProdMulti(nv, X[2], aY, dots))
    # vector arrays (:1700-2690)
    op("N_VLinearSumVectorArray", 24 * nv, lambda: api.LinearSumVectorArray(nv, a, aX, b, aY, aZ))
    op("N_VScaleVectorArray", 16 * nv, lambda: api.ScaleVectorArray(nv, cs8, aX, aZ))
    op("N_VConstVectorArray", 8 * nv, lambda: api.ConstVectorArray(nv, 0.5, aZ))
    op("N_VWrmsNormVectorArray", 16 * nv, lambda: api.WrmsNormVectorArray(nv, aX, aW, nrm))
    op("N_VWrmsNormMaskVectorArray", 8 * (2 * nv + 1), lambda: api.WrmsNormMaskVectorArray(nv, aX, aW, ID, nrm))
    op("N_VScaleAddMultiVectorArray", 8 * (nv + 2 * nv * ns),
       lambda: api.ScaleAddMultiVectorArray(nv, ns, c4, aX, aYY, aZZ))
    op("N_VLinearCombinationVectorArray", 8 * (nv * ns + nv),
       lambda: api.LinearCombinationVectorArray(nv, ns, c4, aYY, aZ))
    # local reductions (no communication even on a distributed vector)
    op("N_VDotProdLocal", 16, lambda: res.__setitem__("dotl", api.DotProdLocal(X[4], Y[4])))
    op("N_VMaxNormLocal", 8, lambda: res.__setitem__("maxl", api.MaxNormLocal(X[5])))
    op("N_VWSqrSumLocal", 16, lambda: res.__setitem__("wsql", api.WSqrSumLocal(X[6], W)))
    op("N_VDotProdMultiLocal", 8 * (nv + 1), lambda: api.DotProdMultiLocal(nv, X[7], aY, dots))
    # the step's result: a checksum of an output vector (read back by the host)
    op("N_VWrmsNorm(result)", 16, lambda: res.__setitem__("result", api.WrmsNorm(Z[1], W)))
    keep = (c8, c8_one, c4, cs8, aX, aY, aZ, aYS, aW, aYY, aZZ, dots, nrm)
    return s, res, keep


def fill_inputs(rng, n):
    import numpy as np

    def pm(lo, hi):
        return rng.uniform(lo, hi, n) * (rng.integers(0, 2, n) * 2 - 1)

    return pm, np


def alloc_vectors(newvec, nv=NVECS, ns=NSUMS):
    return {
        "X": [newvec() for _ in range(nv)], "Y": [newvec() for _ in range(nv)], "Z": [newvec() for _ in range(nv)],
        "S": newvec(), "T": newvec(), "W": newvec(), "ID": newvec(), "CN": newvec(),
        "YY": [[newvec() for _ in range(nv)] for _ in range(ns)],
        "ZZ": [[newvec() for _ in range(nv)] for _ in range(ns)],
    }


def all_handles(vec):
    out = list(vec["X"]) + list(vec["Y"]) + list(vec["Z"]) + [vec["S"], vec["T"], vec["W"], vec["ID"], vec["CN"]]
    for row in vec["YY"] + vec["ZZ"]:
        out += row
    return out


# --------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# --------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # median of the samples under load (upper half: idle samples bracket the region)
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


class LegSampler:
    """nvidia-smi over ALL GPUs of the box while one integrator leg runs (rank 0 only): per GPU
    the median / minimum SM clock, the peak power draw and the throttle reasons seen -- an
    integrator leg advances at the pace of its slowest rank, so one capped GPU shows here."""
    Q = "index," + ClockSampler.Q

    def __init__(self):
        self.rows, self.proc = [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self, ngpus):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        per = {}
        for r in self.rows:
            try:
                g, sm, pw = int(r[0]), float(r[1]), float(r[3])
            except (ValueError, IndexError):
                continue
            if g >= ngpus:
                continue
            d = per.setdefault(g, {"sm": [], "pw": 0.0, "reasons": set()})
            d["sm"].append(sm)
            d["pw"] = max(d["pw"], pw)
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if val.lower().startswith("active"):
                    d["reasons"].add(name)
        out = {}
        for g, d in sorted(per.items()):
            sm = sorted(d["sm"])
            out[f"gpu{g}"] = {"sm_mhz_median": sm[len(sm) // 2], "sm_mhz_min": sm[0], "power_w_max": d["pw"],
                              "reasons": sorted(d["reasons"]), "samples": len(sm)}
        return out


# --------------------------------------------------------------------------
# reference CPU arm (the one place bench.py executes oracle/_ref)
# --------------------------------------------------------------------------
def load_reference():
    so = ROOT / "oracle" / "_ref" / "lib" / "libsundials_ref.so"
    if not so.exists():
        raise FileNotFoundError(f"{so} missing (build with `make -C oracle ref` where /root/reference exists)")
    lib = C.CDLL(str(so), mode=C.RTLD_GLOBAL)
    lib.SUNContext_Create.restype, lib.SUNContext_Create.argtypes = C.c_int, [C.c_int, C.POINTER(C.c_void_p)]
    lib.N_VNew_OpenMP.restype, lib.N_VNew_OpenMP.argtypes = C.c_void_p, [C.c_int64, C.c_int, C.c_void_p]
    lib.N_VNew_Serial.restype, lib.N_VNew_Serial.argtypes = C.c_void_p, [C.c_int64, C.c_void_p]
    lib.N_VEnableFusedOps_OpenMP.restype, lib.N_VEnableFusedOps_OpenMP.argtypes = C.c_int, [C.c_void_p, C.c_int]
    lib.N_VEnableFusedOps_Serial.restype, lib.N_VEnableFusedOps_Serial.argtypes = C.c_int, [C.c_void_p, C.c_int]
    lib.N_VGetArrayPointer.restype, lib.N_VGetArrayPointer.argtypes = C.POINTER(C.c_double), [C.c_void_p]
    return lib


def run_reference_suite(log2n: int, steps: int, warmup: int, threads: int, budget_s: float = 150.0):
    """Time the reference's own CPU vector on the suite.  threads > 1: nvector_openmp,
    threads == 1: nvector_serial.  Returns (GB/s, ms_per_step, sample description)."""
    import numpy as np

    from sundials_b200.plugin import Api

    lib = load_reference()
    ctx = C.c_void_p()
    assert lib.SUNContext_Create(0, C.byref(ctx)) == 0
    api = Api(lib, "")
    n = 1 << log2n
    rng = np.random.default_rng(1234)

    def newvec():
        if threads > 1:
            v = lib.N_VNew_OpenMP(n, threads, ctx)
            lib.N_VEnableFusedOps_OpenMP(v, 1)
        else:
            v = lib.N_VNew_Serial(n, ctx)
            lib.N_VEnableFusedOps_Serial(v, 1)
        return v

    vec = alloc_vectors(newvec)
    _init_values(vec, lambda v: np.ctypeslib.as_array(lib.N_VGetArrayPointer(v), shape=(n,)), rng, n)
    suite, res, _keep = make_suite(api, vec)
    bytes_per_step = sum(b for _, b, _ in suite) * n

    def step():
        for _, _, fn in suite:
            fn()

    t0 = time.perf_counter()
    step()
    first = time.perf_counter() - t0
    # bound the run: fewer timed steps if the box is slow (never fewer than 1)
    warmup = min(warmup, max(0, int(budget_s * 0.2 / first) - 1))
    steps_run = max(1, min(steps, int(budget_s * 0.8 / first)))
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps_run):
        step()
    dt = (time.perf_counter() - t0) / steps_run
    kind = f"nvector_openmp({threads} threads)" if threads > 1 else "nvector_serial(1 core)"
    sample = (f"{kind}, same {len(suite)}-op suite, fused ops enabled, length 2^{log2n} per vector "
              f"(bounded sample), {steps_run} timed steps after {warmup + 1} warm-up")
    return bytes_per_step / dt / 1e9, dt * 1e3, sample, steps_run


def _init_values(vec, host_view, rng, n):
    """Fill every vector with benign seeded data (inputs away from 0, 0/1 mask,
    constraints in {-2..2}) -- mirrors N_VRand / N_VRandZeroOne / N_VRandConstraints
    of the reference benchmark but with a fixed seed."""
    import numpy as np

    for v in all_handles(vec):
        a = host_view(v)
        a[...] = rng.uniform(0.5, 1.5, n) * (rng.integers(0, 2, n) * 2 - 1)
    host_view(vec["W"])[...] = rng.uniform(0.5, 1.5, n)
    host_view(vec["ID"])[...] = rng.integers(0, 2, n).astype(np.float64)
    host_view(vec["CN"])[...] = rng.integers(-2, 3, n).astype(np.float64)


def run_diffusion(ctx, world, args):
    """Bounded solve of the re-hosted ARKODE diffusion_2D benchmark (apps/diffusion_2D) on the
    bench's own context: nx x (ny * world) mesh, strips in y, dx = dy fixed (yu = world), so
    the per-GPU work is the BASELINE config (8192^2 per GPU, DIRK-3 + PCG(20) + Jacobi)."""
    sys.path.insert(0, str(ROOT / "apps" / "diffusion_2D"))
    import run as app

    n = args.diffusion_n
    st = app.run(ctx, nx=n, ny=n * world, yu=float(world), tf=args.diffusion_tf, nout=1, output=0)
    ev = st["evolve_seconds"]
    return {
        "workload": f"benchmarks/diffusion_2D re-host: {n} x {n * world} mesh ({n}^2 per GPU, strips in y, "
                    f"dx = dy = 1/{n - 1}), ARKODE DIRK order 3, PCG liniters 20, Jacobi, rtol 1e-5 atol 1e-10, "
                    f"tf = {args.diffusion_tf} (bounded: the full tf = 1 run is ~1e4 x longer)",
        "solve_s": round(ev, 4), "steps": st["nst"], "step_attempts": st["nst_a"], "ls_iters": st["nli"],
        "rhs_evals": st["rhs_calls"], "ms_per_ls_iter": round(ev / max(st["nli"], 1) * 1e3, 4),
        "ns_per_node_per_ls_iter": round(ev / max(st["nli"], 1) / st["nodes"] * world * 1e9, 5),
        "nodes_per_gpu": st["nodes_loc"], "max_err": st["max_err"],
    }


def run_diffusion_cpu_reference(n=2048, tf=1e-5):
    """The reference's own benchmarks/diffusion_2D (CPU backend, one rank: no MPI in this
    image) on a bounded sample: oracle/_ref/bin/arkode_diffusion_2D_ref."""
    import re

    exe = ROOT / "oracle" / "_ref" / "bin" / "arkode_diffusion_2D_ref"
    if not exe.exists():
        return {"unavailable": f"{exe} missing"}
    t0 = time.perf_counter()
    r = subprocess.run([str(exe), "--nx", str(n), "--ny", str(n), "--tf", str(tf), "--nout", "1"],
                       capture_output=True, text=True, timeout=600)
    dt = time.perf_counter() - t0
    m = re.search(r"^LS iters\s*=\s*(\d+)", r.stdout, re.M)
    nli = int(m.group(1)) if m else 0
    return {"kind": "reference", "cores": 1, "sample": f"{n}^2 mesh, tf = {tf}, whole program wall time",
            "wall_s": round(dt, 3), "ls_iters": nli,
            "ns_per_node_per_ls_iter": round(dt / max(nli, 1) / (n * n) * 1e9, 3)}


def run_ar3d(ctx, world, args):
    """Bounded solve of the re-hosted benchmarks/advection_reaction_3D (apps/advection_reaction_3D)
    on the bench's own context: BASELINE config 5 -- npts^3 mesh (512^3 = 4.0e8 unknowns), slabs
    in x over the ranks (STRONG scaling: the mesh is fixed), ARKODE IMEX-ARK order 3, Newton +
    SPGMR with the reaction-block preconditioner, fused vector ops on."""
    sys.path.insert(0, str(ROOT / "apps" / "advection_reaction_3D"))
    import importlib.util

    spec = importlib.util.spec_from_file_location("ar3d_run", ROOT / "apps" / "advection_reaction_3D" / "run.py")
    app = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(app)
    n = args.ar3d_npts
    st = app.run(ctx, npts=n, method="ARK-IMEX", nls="newton", fused=1, tf=args.ar3d_tf, nout=1, output=0)
    ev = st["evolve_seconds"]
    return {
        "workload": f"benchmarks/advection_reaction_3D re-host: {n}^3 mesh x 3 species = {st['neq']} unknowns, "
                    f"slabs in x over {world} GPU(s) (strong scaling), ARKODE IMEX-ARK order 3, Newton + SPGMR + "
                    f"block preconditioner, rtol 1e-6 atol 1e-9, fused ops, tf = {args.ar3d_tf} (bounded; the "
                    f"benchmark's default is tf = 10)",
        "scaling": "strong", "solve_s": round(ev, 4), "steps": st["nst"], "step_attempts": st["nst_a"],
        "fe_evals": st["nfe"], "fi_evals": st["nfi"], "nls_iters": st["nni"], "ls_iters": st["nli"],
        "prec_solves": st["npsol"], "ms_per_step": round(ev / max(st["nst"], 1) * 1e3, 3),
        "ns_per_unknown_per_step": round(ev / max(st["nst"], 1) / st["neq"] * 1e9, 6),
        "unknowns_per_gpu": st["neq_loc"], "urms": st["urms"], "vrms": st["vrms"], "wrms": st["wrms"],
    }


def run_ar3d_cpu_reference(n=64, tf=0.05):
    """The reference's own benchmarks/advection_reaction_3D (RAJA sequential backend, one rank)
    on a bounded sample: oracle/_ref/bin/advection_reaction_3D_ref."""
    import re
    import tempfile

    exe = ROOT / "oracle" / "_ref" / "bin" / "advection_reaction_3D_ref"
    if not exe.exists():
        return {"unavailable": f"{exe} missing"}
    with tempfile.TemporaryDirectory() as td:
        t0 = time.perf_counter()
        r = subprocess.run([str(exe), "--npts", str(n), "--method", "ARK-IMEX", "--nls", "newton", "--fused", "--tf",
                            str(tf), "--nout", "1", "--dont-save", "--output-dir", td], capture_output=True, text=True,
                           timeout=600, cwd=td)
        dt = time.perf_counter() - t0
    m = re.search(r"Internal solver steps = (\d+)", r.stdout)
    nst = int(m.group(1)) if m else 0
    return {"kind": "reference", "cores": 1, "sample": f"{n}^3 mesh, tf = {tf}, whole program wall time",
            "wall_s": round(dt, 3), "steps": nst,
            "ns_per_unknown_per_step": round(dt / max(nst, 1) / (3 * n ** 3) * 1e9, 3)}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    log2n = min(args.log2n, args.cpu_log2n)
    try:
        gbs, ms, sample, steps_run = run_reference_suite(log2n, args.steps, args.warmup, threads)
    except FileNotFoundError as e:
        print(json.dumps({"impl": "reference", "unavailable": str(e)}))
        return 0
    line = {
        "impl": "reference", "metric": "N_Vector op suite throughput (algorithmic GB/s)", "value": round(gbs, 2),
        "unit": "GB/s", "n_gpus": args.gpus, "steps": steps_run, "warmup": args.warmup,
        "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": round(gbs, 2), "unit": "GB/s", "cores": threads, "kind": "reference",
                         "sample": sample},
        "e2e": {"value": round(gbs, 2), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def workload_config(args):
    return {
        "workload": "benchmarks/nvector performance suite (all ops incl. fused/vector-array), "
                    f"length 2^{args.log2n} per GPU, nvecs={NVECS}, nsums={NSUMS}, fused ops enabled",
        "length_per_gpu": 1 << args.log2n,
        "nvecs": NVECS, "nsums": NSUMS,
        "partition": "contiguous 1-D block per GPU (MPIPlusX pattern); reductions fold the ranks' partials over "
                     "NVLink peer memory inside the reduction kernel (NCCL allreduce as fallback); no other communication",
        "cache": "inputs larger than L2: every op streams >= 128 MiB per operand (126 MB L2), "
                 "91 distinct vectors (11.4 GiB) rotate through the suite",
    }


# --------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------
def b200_arm(args):
    import numpy as np
    import torch

    from sundials_b200 import _lib
    from sundials_b200.plugin import B200Plugin

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: anything native libraries print there
    # (e.g. NCCL's version banner) goes to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        import datetime

        # a mismatched collective must fail in minutes, not hold the GPUs for the default 10
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"),
                                timeout=datetime.timedelta(seconds=180))

    lib = _lib.load()
    P = B200Plugin()
    n = 1 << args.log2n

    # one execution context per rank on the legacy stream (+ NCCL communicator)
    ctx = C.c_void_p()
    _lib.check(lib.b200vec_ctx_create(C.byref(ctx), local_rank, None), "ctx_create")
    if world > 1:
        idbuf = (C.c_ubyte * _lib.UNIQUE_ID_BYTES)()
        if rank == 0:
            _lib.check(lib.b200vec_comm_get_unique_id(idbuf), "comm_get_unique_id")
        t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, src=0)
        idbuf = (C.c_ubyte * _lib.UNIQUE_ID_BYTES).from_buffer_copy(bytes(t.cpu().tolist()))
        _lib.check(lib.b200vec_comm_init(ctx, idbuf, rank, world), "comm_init")

    def newvec():
        v = P.new(n, ctx, P.DEVICE, fused=True)
        if world > 1:
            assert lib.N_VMakeDistributed_B200(v, n * world) == 0
        return v

    vec = alloc_vectors(newvec)
    rng = np.random.default_rng(1234 + rank)
    # initialise through the plugin: pinned host mirror -> N_VCopyToDevice; only the
    # nvecs input vectors X keep their host mirror (the e2e leg re-uploads them)
    keep_host = set(vec["X"]) | {vec["Z"][1]}
    for v in all_handles(vec):
        a = P.host(v, n)
        a[...] = rng.uniform(0.5, 1.5, n) * (rng.integers(0, 2, n) * 2 - 1)
        if v == vec["W"]:
            a[...] = np.abs(a)
        elif v == vec["ID"]:
            a[...] = rng.integers(0, 2, n)
        elif v == vec["CN"]:
            a[...] = rng.integers(-2, 3, n)
        P.to_device(v)
        if v not in keep_host:
            P.drop_host(v)
    suite, res, _keep = make_suite(P, vec)
    bytes_per_step = sum(b for _, b, _ in suite) * n

    def step():
        for _, _, fn in suite:
            fn()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        """k steps bracketed by barrier+sync, CUDA events on the launching (legacy) stream."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / k

    # ---- value: operands resident in HBM
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lib.b200vec_ctx_set_tuning(ctx, b"count_launches", 1)
    ms_step = timed(step, args.steps)
    launches = int(lib.b200vec_ctx_launch_count(ctx))
    lib.b200vec_ctx_set_tuning(ctx, b"count_launches", 0)
    result_value = res.get("result")

    # ---- e2e: H2D of the nvecs input vectors + suite + D2H of a result vector and scalars
    h2d = NVECS * n * 8
    d2h = n * 8 + 8 * 24

    def step_e2e():
        for v in vec["X"]:
            P.to_device(v)      # pinned host mirror -> HBM (cudaMemcpyAsync + sync)
        step()
        P.from_device(vec["Z"][1])

    for _ in range(min(3, args.warmup)):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    value = world * bytes_per_step / (ms_step * 1e-3) / 1e9
    e2e_value = world * bytes_per_step / (ms_e2e * 1e-3) / 1e9

    # ---- per-op timings: each op alone, CUDA events.  EVERY rank runs this loop --
    # the reducing ops of a distributed vector are collectives (SPMD) -- rank 0 reports
    per_op = {}
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    reps = 10
    for name, bpe, fn in suite:
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        per_op[name] = {"us": round(us, 2), "GBs": round(bpe * n / us / 1e3, 1),
                        "frac_of_peak": round(bpe * n / us / 1e3 / peak, 3)}

    # reductions: the API-level time above includes the host round trip that
    # returning a scalar requires; also time the kernels alone (C ABI, async)
    dptr = lib.N_VGetDeviceArrayPointer_B200
    X, Y, Z = vec["X"], vec["Y"], vec["Z"]
    W, ID, CN = vec["W"], vec["ID"], vec["CN"]

    def tab(vs):
        return (C.c_void_p * len(vs))(*[dptr(v) for v in vs])

    tX, tY = tab(X), tab(Y)
    kernel_only = {
        "N_VDotProd": (16, lambda: lib.b200vec_dot_prod(ctx, dptr(X[1]), dptr(Y[1]), n, None)),
        "N_VMaxNorm": (8, lambda: lib.b200vec_max_norm(ctx, dptr(X[2]), n, None)),
        "N_VWrmsNorm": (16, lambda: lib.b200vec_wsqr_sum(ctx, dptr(X[3]), dptr(W), n, None)),
        "N_VWrmsNormMask": (24, lambda: lib.b200vec_wsqr_sum_mask(ctx, dptr(X[4]), dptr(W), dptr(ID), n, None)),
        "N_VMin": (8, lambda: lib.b200vec_min(ctx, dptr(X[5]), n, None)),
        "N_VL1Norm": (8, lambda: lib.b200vec_l1_norm(ctx, dptr(X[7]), n, None)),
        "N_VInvTest": (16, lambda: lib.b200vec_inv_test(ctx, dptr(X[1]), dptr(Z[1]), n, None)),
        "N_VConstrMask": (24, lambda: lib.b200vec_constr_mask(ctx, dptr(CN), dptr(X[2]), dptr(Z[2]), n, None)),
        "N_VMinQuotient": (16, lambda: lib.b200vec_min_quotient(ctx, dptr(X[3]), dptr(Y[3]), n, None)),
        "N_VDotProdMulti": (8 * (NVECS + 1), lambda: lib.b200vec_dot_prod_multi(ctx, NVECS, dptr(X[2]), tY, n, None)),
        "N_VWrmsNormVectorArray": (16 * NVECS,
                                   lambda: lib.b200vec_wsqr_sum_vector_array(ctx, NVECS, tX, tY, None, n, None)),
        "N_VWrmsNormMaskVectorArray": (8 * (2 * NVECS + 1),
                                       lambda: lib.b200vec_wsqr_sum_vector_array(ctx, NVECS, tX, tY, dptr(ID), n, None)),
    }
    for name, (bpe, fn) in kernel_only.items():
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        per_op[name].update({"kernel_us": round(us, 2), "kernel_GBs": round(bpe * n / us / 1e3, 1),
                             "kernel_frac_of_peak": round(bpe * n / us / 1e3 / peak, 3)})

    # dominant kernel by share of the step (profiles/r01_launches_summary.md: 28 %):
    # k_scaleadd_rows<4>, launched 3x per step (N_VScaleAddMulti-1/-2, 136 B/elt, and
    # N_VScaleAddMultiVectorArray, 576 B/elt).  achieved = algorithmic bytes of
    # those launches / their CUDA-event time, measured live above.
    dom_ops = ["N_VScaleAddMulti-1", "N_VScaleAddMulti-2", "N_VScaleAddMultiVectorArray"]
    dom_bytes = {name: bpe * n for name, bpe, _ in suite if name in dom_ops}
    dom_us = sum(per_op[k]["us"] for k in dom_ops)
    dom_gbs = sum(dom_bytes.values()) / dom_us / 1e3
    roofline = {"bound": "hbm", "kernel": "k_scaleadd_rows<4> (N_VScaleAddMulti, N_VScaleAddMultiVectorArray)",
                "achieved": round(dom_gbs, 1), "peak": peak, "unit": "GB/s", "frac": round(dom_gbs / peak, 4),
                "traffic": None, "peak_source": peak_src,
                "launches_per_step": 3,
                "algorithmic_bytes_per_launch": {k: dom_bytes[k] for k in dom_ops},
                "avg_launch_us": round(dom_us / 3, 2),
                "share_of_step": round(dom_us / (ms_step * 1e3), 4),
                "suite_frac": round(value / world / peak, 4)}
    prof = ROOT / "profiles" / "scaleadd_traffic.json"
    if prof.exists():
        try:
            t = json.loads(prof.read_text())
            # ncu --set full capture of the N_VScaleAddMulti launch (nv=8, n=2^24): dram read+write per launch
            roofline["traffic"] = t.get("dram_bytes_per_launch")
            roofline["traffic_launch"] = t.get("kernel")
        except Exception:
            pass

    # ---- ARKODE diffusion_2D solve time (second half of BASELINE's metric); collective
    diffusion = None
    if not args.no_diffusion:
        leg = LegSampler().start() if rank == 0 else None
        try:
            diffusion = run_diffusion(ctx, world, args)
        except Exception as e:  # reported, never required for the op-suite line
            diffusion = {"unavailable": f"{type(e).__name__}: {e}"}
        if leg is not None:
            diffusion["clocks"] = leg.stop(world)

    # ---- ARKODE advection_reaction_3D (BASELINE config 5); collective
    ar3d = None
    if not args.no_ar3d:
        leg = LegSampler().start() if rank == 0 else None
        try:
            ar3d = run_ar3d(ctx, world, args)
        except Exception as e:
            ar3d = {"unavailable": f"{type(e).__name__}: {e}"}
        if leg is not None:
            ar3d["clocks"] = leg.stop(world)

    # ---- Krylov Gram-Schmidt built on the ops: the reference's unmodified
    # SUNClassicalGS / SUNModifiedGS on a basis of this vector; collective
    gs = None
    if not args.no_gs:
        try:
            sys.path.insert(0, str(ROOT / "tools"))
            import gs_bench

            gs = gs_bench.run(args.log2n, maxl=5, reps=5, cpu_log2n=args.cpu_log2n, with_ref_cuda=False,
                              with_cpu=(world == 1 and not args.no_cpu_baseline), b200_ctx=ctx, rank=rank, world=world)
            for k in ("classical", "modified"):
                gs["b200"][k]["cycle_frac_of_peak"] = round(gs["b200"][k]["cycle_GBs"] / peak, 3)
        except Exception as e:
            gs = {"unavailable": f"{type(e).__name__}: {e}"}

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    if ar3d is not None and "solve_s" in ar3d and world == 1 and not args.no_cpu_baseline:
        try:
            ar3d["cpu_reference"] = run_ar3d_cpu_reference(tf=args.ar3d_tf)
        except Exception as e:
            ar3d["cpu_reference"] = {"unavailable": str(e)}

    if diffusion is not None and "solve_s" in diffusion and world == 1 and not args.no_cpu_baseline:
        try:
            diffusion["cpu_reference"] = run_diffusion_cpu_reference()
        except Exception as e:
            diffusion["cpu_reference"] = {"unavailable": str(e)}

    # ---- CPU baseline on this box's host cores (bounded sample)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            threads = os.cpu_count() or 1
            g_omp, _, sample_omp, _ = run_reference_suite(args.cpu_log2n, 2, 1, threads, budget_s=25.0)
            g_ser, _, sample_ser, _ = run_reference_suite(min(args.cpu_log2n, 20), 1, 0, 1, budget_s=15.0)
            cpu = {"value": round(g_omp, 2), "unit": "GB/s", "cores": threads, "kind": "reference",
                   "sample": sample_omp, "serial_1core_GBs": round(g_ser, 2), "serial_sample": sample_ser}
        except Exception as e:  # the baseline is reported, never required
            cpu = {"value": None, "unit": "GB/s", "cores": os.cpu_count(), "kind": "reference",
                   "sample": f"unavailable: {e}"}

    cfg = workload_config(args)
    lib.b200vec_comm_transport.restype = C.c_char_p
    cfg["reduction_transport"] = lib.b200vec_comm_transport(ctx).decode()
    line = {
        "metric": "N_Vector op suite throughput (algorithmic GB/s)", "value": round(value, 1), "unit": "GB/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "e2e": {"value": round(e2e_value, 1), "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": round(ms_e2e, 4)},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "clocks": clocks,
        "diffusion_2D": diffusion,
        "advection_reaction_3D": ar3d,
        "gram_schmidt": gs,
        "per_op": per_op,
        "result_checksum": result_value,
        "ops_per_step": len(suite),
    }
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2n", type=int, default=24, help="log2 of the vector length per GPU")
    ap.add_argument("--cpu-log2n", type=int, default=22, help="length of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-diffusion", action="store_true", help="skip the ARKODE diffusion_2D leg")
    ap.add_argument("--diffusion-n", type=int, default=8192, help="mesh points per GPU in x and y")
    ap.add_argument("--diffusion-tf", type=float, default=1e-4)
    ap.add_argument("--no-ar3d", action="store_true", help="skip the ARKODE advection_reaction_3D leg")
    ap.add_argument("--ar3d-npts", type=int, default=512, help="GLOBAL mesh points per direction")
    ap.add_argument("--ar3d-tf", type=float, default=0.05)
    ap.add_argument("--no-gs", action="store_true", help="skip the Gram-Schmidt leg")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return reference_arm(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun (the driver launches torchrun itself)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", __file__] + sys.argv[1:]
        return subprocess.call(cmd)
    return b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
