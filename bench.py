#!/usr/bin/env python
"""bench.py -- N_Vector op throughput on B200 (BASELINE.json metric).

A *step* is one pass of the reference's N_Vector performance suite
(benchmarks/nvector/test_nvector_performance.c: every standard, reduction, fused
and vector-array op, 16 N_VLinearSum cases, 4 N_VScale cases, ...) over vectors
of length 2^LOG2N per GPU with nvecs=8, nsums=4, fused ops ENABLED.  The suite is
a C driver (apps/nvector_perf/nvector_perf.c, a re-host of the reference harness)
that reaches every op through `v->ops->nv...` -- the SUNDIALS ops table, the
drop-in boundary -- exactly as the integrators do; the SAME compiled driver runs
both arms (NVECTOR_B200 here, the reference's nvector_openmp under --impl reference).
Throughput is ALGORITHMIC bytes (SURVEY.md section 8d byte model: each distinct
operand read once, each output written once) per second, summed over the suite.

  value   suite GB/s with all operands resident in HBM (CUDA events, max over ranks)
  e2e     same suite, every step's nvecs input vectors uploaded from pinned host memory
          (N_VCopyToDeviceAsync_B200 on the context's copy stream, double-buffered: the upload
          of step k+1 overlaps the kernels of step k) and a result vector + scalars read back
  roofline  the dominant kernel by share of the step (k_scaleadd_rows<4>) timed
          with CUDA events, vs the measured HBM copy peak
  cpu_baseline  the reference's own nvector_openmp (all host threads) and
          nvector_serial (1 core) on a bounded sample of the same suite
  legs    (printed LAST) compact record of the integrator / Gram-Schmidt / sweep legs
  --impl reference : the reference CPU implementation alone (driver's baseline arm)
  --sweep          : BASELINE configs[2], lengths 2^16 .. 2^30 per GPU, CPU columns beside

Multi-GPU (torchrun, one rank per GPU): weak scaling, each rank owns the local
block of a contiguous 1-D partition (MPIPlusX pattern); streaming ops need no
communication, every reduction folds the ranks' partials inside its kernel over
NVLink peer memory (ncclAllReduce as the fallback transport).  Before timing, a
distributed exact-answer parity block runs on BOTH transports (dist_parity).
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
V = C.c_void_p


# --------------------------------------------------------------------------
# the suite: C driver over the ops table (apps/nvector_perf)
# --------------------------------------------------------------------------
def load_perf():
    so = ROOT / "apps" / "nvector_perf" / "_build" / "libnvector_perf.so"
    if not so.exists():
        raise FileNotFoundError(f"{so} missing (run `make -C apps/nvector_perf` where the SUNDIALS headers exist)")
    lib = C.CDLL(str(so))
    Vp = C.POINTER(V)
    lib.nvperf_create.restype, lib.nvperf_create.argtypes = V, [Vp, Vp, Vp, V, V, V, V, V, Vp, Vp, C.c_int, C.c_int]
    lib.nvperf_destroy.restype, lib.nvperf_destroy.argtypes = None, [V]
    lib.nvperf_num_ops.restype = C.c_int
    lib.nvperf_op_name.restype, lib.nvperf_op_name.argtypes = C.c_char_p, [C.c_int]
    lib.nvperf_op_bytes_per_elt.restype, lib.nvperf_op_bytes_per_elt.argtypes = C.c_double, [V, C.c_int]
    lib.nvperf_op_returns_scalar.restype, lib.nvperf_op_returns_scalar.argtypes = C.c_int, [C.c_int]
    lib.nvperf_run_op.restype, lib.nvperf_run_op.argtypes = None, [V, C.c_int, C.c_int]
    lib.nvperf_run_step.restype, lib.nvperf_run_step.argtypes = None, [V, C.c_int]
    lib.nvperf_result.restype, lib.nvperf_result.argtypes = C.c_double, [V, C.c_int]
    lib.nvperf_error.restype, lib.nvperf_error.argtypes = C.c_int, [V]
    return lib


class Suite:
    """one nvperf suite over a dict of vector handles (see alloc_vectors)"""

    def __init__(self, perf, vec, X=None):
        self.perf = perf
        X = X or vec["X"]
        nv, ns = len(X), len(vec["YY"])
        arr = lambda hs: (V * len(hs))(*hs)  # noqa: E731
        self._keep = (arr(X), arr(vec["Y"]), arr(vec["Z"]), arr([h for row in vec["YY"] for h in row]),
                      arr([h for row in vec["ZZ"] for h in row]))
        self.h = perf.nvperf_create(self._keep[0], self._keep[1], self._keep[2], vec["S"], vec["T"], vec["W"], vec["ID"],
                                    vec["CN"], self._keep[3], self._keep[4], nv, ns)
        if not self.h:
            raise RuntimeError("nvperf_create failed (fused ops not enabled on the vector?)")
        self.nops = perf.nvperf_num_ops()
        self.names = [perf.nvperf_op_name(i).decode() for i in range(self.nops)]
        self.bpe = [perf.nvperf_op_bytes_per_elt(self.h, i) for i in range(self.nops)]
        self.scalar = [bool(perf.nvperf_op_returns_scalar(i)) for i in range(self.nops)]
        self.bytes_per_elt_step = sum(self.bpe)

    def step(self, k=1):
        self.perf.nvperf_run_step(self.h, k)

    def op(self, i, reps=1):
        self.perf.nvperf_run_op(self.h, i, reps)

    def result(self, name):
        return self.perf.nvperf_result(self.h, self.names.index(name))

    def check(self):
        e = self.perf.nvperf_error(self.h)
        if e:
            raise RuntimeError(f"a fused op of the suite returned error {e}")


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


def base_arrays(n, seed):
    """Seeded synthetic data, mirroring N_VRand / N_VRandZeroOne / N_VRandConstraints of the reference
    benchmark (test_nvector_performance.c:2751-2800) with a FIXED seed: signed values away from 0
    (divisors), positive weights, a 0/1 mask, constraints in {-2..2}.  Vector i of the suite holds
    scale(i) * u -- cheap to build for 91 vectors on either arm, and identical on both arms."""
    import numpy as np

    rng = np.random.default_rng(seed)
    u = rng.uniform(0.5, 1.5, n) * (rng.integers(0, 2, n) * 2 - 1)
    w = rng.uniform(0.5, 1.5, n)
    idm = rng.integers(0, 2, n).astype(np.float64)
    cn = rng.integers(-2, 3, n).astype(np.float64)
    return u, w, idm, cn


def vec_scale(i):
    return 0.75 + 0.5 * ((i * 0.6180339887498949) % 1.0)


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
# reference libraries.  load_host: the unmodified host framework NVECTOR_B200 plugs into
# (baseline/_ref: core, solvers, integrators).  load_reference: + the reference's CPU vectors
# (oracle/_ref) -- the ONE place bench.py executes oracle/, and only in the CPU legs.
# --------------------------------------------------------------------------
def load_host():
    so = ROOT / "baseline" / "_ref" / "lib" / "libsundials_host.so"
    if not so.exists():
        raise FileNotFoundError(f"{so} missing (build with `make -C baseline` where /root/reference exists)")
    lib = C.CDLL(str(so), mode=C.RTLD_GLOBAL)
    lib.SUNContext_Create.restype, lib.SUNContext_Create.argtypes = C.c_int, [C.c_int, C.POINTER(C.c_void_p)]
    lib.N_VGetArrayPointer.restype, lib.N_VGetArrayPointer.argtypes = C.POINTER(C.c_double), [C.c_void_p]
    lib.N_VDestroy.restype, lib.N_VDestroy.argtypes = None, [C.c_void_p]
    return lib


def load_reference():
    so = ROOT / "oracle" / "_ref" / "lib" / "libsundials_ref.so"
    if not so.exists():
        raise FileNotFoundError(f"{so} missing (build with `make -C oracle ref` where /root/reference exists)")
    load_host()
    lib = C.CDLL(str(so), mode=C.RTLD_GLOBAL)
    lib.SUNContext_Create.restype, lib.SUNContext_Create.argtypes = C.c_int, [C.c_int, C.POINTER(C.c_void_p)]
    lib.N_VNew_OpenMP.restype, lib.N_VNew_OpenMP.argtypes = C.c_void_p, [C.c_int64, C.c_int, C.c_void_p]
    lib.N_VNew_Serial.restype, lib.N_VNew_Serial.argtypes = C.c_void_p, [C.c_int64, C.c_void_p]
    lib.N_VEnableFusedOps_OpenMP.restype, lib.N_VEnableFusedOps_OpenMP.argtypes = C.c_int, [C.c_void_p, C.c_int]
    lib.N_VEnableFusedOps_Serial.restype, lib.N_VEnableFusedOps_Serial.argtypes = C.c_int, [C.c_void_p, C.c_int]
    lib.N_VGetArrayPointer.restype, lib.N_VGetArrayPointer.argtypes = C.POINTER(C.c_double), [C.c_void_p]
    lib.N_VDestroy.restype, lib.N_VDestroy.argtypes = None, [C.c_void_p]
    return lib


def host_mem_available_bytes():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except OSError:
        pass
    return 0


def cpu_vectors(lib, sctx, n, threads, nv=NVECS, ns=NSUMS, seed=1234):
    """the suite's vectors on the reference's own CPU vector: nvector_openmp (threads > 1) or
    nvector_serial, fused ops enabled, the same data as the B200 arm (base_arrays / vec_scale)"""
    import numpy as np

    def newvec():
        if threads > 1:
            v = lib.N_VNew_OpenMP(n, threads, sctx)
            lib.N_VEnableFusedOps_OpenMP(v, 1)
        else:
            v = lib.N_VNew_Serial(n, sctx)
            lib.N_VEnableFusedOps_Serial(v, 1)
        assert v, "reference vector constructor failed (out of host memory?)"
        return v

    vec = alloc_vectors(newvec, nv, ns)
    view = lambda v: np.ctypeslib.as_array(lib.N_VGetArrayPointer(v), shape=(n,))  # noqa: E731
    u, w, idm, cn = base_arrays(n, seed)
    for i, v in enumerate(all_handles(vec)):
        np.multiply(u, vec_scale(i), out=view(v))
    np.multiply(w, vec_scale(1), out=view(vec["W"]))
    view(vec["ID"])[...] = idm
    view(vec["CN"])[...] = cn
    return vec


def free_vectors(lib, vec):
    for v in all_handles(vec):
        lib.N_VDestroy(v)


def run_reference_suite(log2n: int, steps: int, warmup: int, threads: int, budget_s: float = 150.0):
    """Time the reference's own CPU vector on the suite (the same C driver as the B200 arm).
    threads > 1: nvector_openmp, threads == 1: nvector_serial.
    Returns (GB/s, ms_per_step, sample description, steps run, result checksum)."""
    lib = load_reference()
    perf = load_perf()
    sctx = C.c_void_p()
    assert lib.SUNContext_Create(0, C.byref(sctx)) == 0
    n = 1 << log2n
    vec = cpu_vectors(lib, sctx, n, threads)
    suite = Suite(perf, vec)
    bytes_per_step = suite.bytes_per_elt_step * n
    t0 = time.perf_counter()
    suite.step(1)
    first = time.perf_counter() - t0
    # bound the run: fewer timed steps if the box is slow (never fewer than 1)
    warmup = min(warmup, max(0, int(budget_s * 0.2 / first) - 1))
    steps_run = max(1, min(steps, int(budget_s * 0.8 / first)))
    suite.step(warmup)
    t0 = time.perf_counter()
    suite.step(steps_run)
    dt = (time.perf_counter() - t0) / steps_run
    suite.check()
    checksum = suite.result("N_VWrmsNorm(result)")
    kind = f"nvector_openmp({threads} threads)" if threads > 1 else "nvector_serial(1 core)"
    sample = (f"{kind}, same {suite.nops}-op suite (C driver over the ops table), fused ops enabled, length "
              f"2^{log2n} per vector, {steps_run} timed steps after {warmup + 1} warm-up")
    free_vectors(lib, vec)
    return bytes_per_step / dt / 1e9, dt * 1e3, sample, steps_run, checksum


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


def run_ar3d(ctx, world, args, nls="newton", tf=None):
    """Bounded solve of the re-hosted benchmarks/advection_reaction_3D (apps/advection_reaction_3D)
    on the bench's own context: BASELINE config 5 -- npts^3 mesh (512^3 = 4.0e8 unknowns), slabs
    in x over the ranks (STRONG scaling: the mesh is fixed), ARKODE IMEX-ARK order 3, fused vector ops on,
    with either nonlinear solver the config names: Newton + SPGMR with the reaction-block preconditioner,
    or the Anderson-accelerated fixed point (3 vectors)."""
    sys.path.insert(0, str(ROOT / "apps" / "advection_reaction_3D"))
    import importlib.util

    spec = importlib.util.spec_from_file_location("ar3d_run", ROOT / "apps" / "advection_reaction_3D" / "run.py")
    app = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(app)
    n = args.ar3d_npts
    tf = args.ar3d_tf if tf is None else tf
    st = app.run(ctx, npts=n, method="ARK-IMEX", nls=nls, fused=1, tf=tf, nout=1, output=0)
    ev = st["evolve_seconds"]
    how = "Newton + SPGMR + block preconditioner" if nls == "newton" else "Anderson-accelerated fixed point (m = 3)"
    return {
        "workload": f"benchmarks/advection_reaction_3D re-host: {n}^3 mesh x 3 species = {st['neq']} unknowns, "
                    f"slabs in x over {world} GPU(s) (strong scaling), ARKODE IMEX-ARK order 3, {how}, "
                    f"rtol 1e-6 atol 1e-9, fused ops, tf = {tf} (bounded; the "
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


def run_cvdiurnal():
    """BASELINE config 2: the reference's examples/cvode/serial/cvDiurnal_kry.c, unmodified, on
    NVECTOR_B200 (host-coherent pinned mode: the example's RHS runs on the host through
    N_VGetArrayPointer) against the same source on nvector_serial.  N = 200 unknowns: launch-latency
    bound by construction -- reported honestly ("even though slower")."""
    b200 = ROOT / "baseline" / "_ref" / "bin" / "cvDiurnal_kry_b200"
    ser = ROOT / "oracle" / "_ref" / "bin" / "cvDiurnal_kry_serial"
    if not b200.exists():
        return {"unavailable": f"{b200} missing"}

    import re

    rep = {}

    def best(exe, reps):
        t, out = 1e30, ""
        for _ in range(reps):
            t0 = time.perf_counter()
            r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600,
                               env=dict(os.environ, B200VEC_REPORT="1"))
            dt = time.perf_counter() - t0
            if dt < t:
                t, out = dt, r.stdout
                m = re.search(r"\[b200vec\] report: (\{.*\})", r.stderr)
                if m:
                    rep.update(json.loads(m.group(1)))
        return t, out

    tb, ob = best(b200, 2)
    leg = {"workload": "CVODE cvDiurnal_kry (2-species diurnal kinetics, BDF + SPGMR, N = 200), whole-program wall "
                       "time, best of 2; NVECTOR_B200 in its host-coherent pinned mode (every op synchronises)",
           "b200_pinned_s": round(tb, 3)}
    if rep:  # time inside the vector's context (without CUDA initialisation / process start-up) and its launches
        leg["b200_ctx_lifetime_s"] = rep.get("ctx_lifetime_s")
        leg["b200_kernel_launches"] = rep.get("kernel_launches")
        if rep.get("kernel_launches"):
            leg["us_per_vector_op"] = round(rep["ctx_lifetime_s"] / rep["kernel_launches"] * 1e6, 2)
    gold = ROOT / "tests" / "golden" / "examples" / "cvDiurnal_kry.out"
    if gold.exists():
        leg["stdout_identical_to_serial_golden"] = (ob == gold.read_text())
    if ser.exists():
        ts, os_ = best(ser, 3)
        leg["serial_1core_s"] = round(ts, 3)
        leg["stdout_identical_to_serial_run"] = (ob == os_)
    return leg


def dist_parity(P, lib, ctx, rank, world, dist, torch):
    """world > 1: exact-answer checks through N_V*_B200 on DISTRIBUTED vectors, on both transports
    (peer-memory fold inside the reduction kernel, and ncclAllReduce), bit-equal between the transports
    and across the ranks.  Known answers in the style of the reference's MPI vector tests
    (test/unit_tests/nvector/test_nvector.c: global-length answers; Test_N_VDotProdMultiAllReduce
    :5663-5794): exactly representable data, one sentinel per rank for max / min, a zero on the last
    rank for N_VInvTest, a 3-wide N_VDotProdMulti, the fused linear-combination + norm."""
    import numpy as np

    fails, checks, bits = [], 0, {}
    transports = []
    for p2p in (1, 0):
        lib.b200vec_ctx_set_tuning(ctx, b"p2p", p2p)
        lib.b200vec_comm_transport.restype = C.c_char_p
        tname = lib.b200vec_comm_transport(ctx).decode()
        transports.append(tname)
        got = []
        for n in (1000, 1 << 20):       # exact-order path and the tree path
            ng = n * world

            def mk(val):
                v = P.new(n, ctx, P.DEVICE, fused=True)
                assert lib.N_VMakeDistributed_B200(v, ng) == 0
                P.Const(val, v)
                return v

            x, y, z = mk(2.0), mk(0.5), mk(0.0)
            exp = [("N_VGetLength", float(P.GetLength(x)), float(ng)),
                   ("N_VDotProd", P.DotProd(x, y), float(ng)),
                   ("N_VL1Norm", P.L1Norm(x), 2.0 * ng),
                   ("N_VWrmsNorm", P.WrmsNorm(x, y), 1.0),
                   ("N_VWL2Norm", P.WL2Norm(x, y), float(np.sqrt(float(ng)))),
                   ("N_VDotProdLocal", P.DotProdLocal(x, y), float(n)),
                   ("N_VInvTest(no zero)", float(P.InvTest(x, z)), 1.0),
                   ("N_VMinQuotient", P.MinQuotient(x, y), 4.0)]
            d = (C.c_double * 3)()
            Y3 = P.varray([x, y, x])
            assert P.DotProdMulti(3, x, Y3, d) == 0
            exp += [("N_VDotProdMulti[0]", d[0], 4.0 * ng), ("N_VDotProdMulti[1]", d[1], float(ng)),
                    ("N_VDotProdMulti[2]", d[2], 4.0 * ng)]
            sq = C.c_double()
            lib.N_VLinearCombinationSqNorm_B200.restype = C.c_int
            lib.N_VLinearCombinationSqNorm_B200.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(V), V,
                                                            C.POINTER(C.c_double)]
            assert lib.N_VLinearCombinationSqNorm_B200(3, P.coefs([1.0, 2.0, -1.0]), P.varray([z, y, x]), z,
                                                       C.byref(sq)) == 0      # z = 1/2 + 1 - 2 = -1/2
            exp.append(("N_VLinearCombinationSqNorm", sq.value, 0.25 * ng))
            # sentinels: rank r plants +(10 + r) and -(20 + r); the last rank plants a zero
            h = P.host(x, n)
            P.from_device(x)
            h[7 + rank] = 10.0 + rank
            h[n - 9 - rank] = -(20.0 + rank)
            P.to_device(x)
            exp += [("N_VMaxNorm", P.MaxNorm(x), 20.0 + world - 1), ("N_VMin", P.Min(x), -(20.0 + world - 1)),
                    ("N_VMaxNormLocal", P.MaxNormLocal(x), 20.0 + rank)]
            if rank == world - 1:
                h[n // 2] = 0.0
                P.to_device(x)
            exp.append(("N_VInvTest(zero on last rank)", float(P.InvTest(x, z)), 0.0))
            for name, g, w in exp:
                checks += 1
                got.append(g)
                if g != w:
                    fails.append(f"{tname} n={n} {name}: got {g!r} want {w!r}")
            for v in (x, y, z):
                P.Destroy(v)
        bits[tname] = np.array(got, dtype=np.float64)
    lib.b200vec_ctx_set_tuning(ctx, b"p2p", 1)
    # identical bits on both transports and on every rank
    a = bits[transports[0]]
    if len(set(transports)) == 2 and not np.array_equal(a.view(np.uint64), bits[transports[1]].view(np.uint64)):
        fails.append("results differ between the transports")
    skip = {i for i in range(len(a)) if i % 16 in (5, 14)}   # the *Local values differ by rank by design
    t = torch.tensor(a, dtype=torch.float64, device="cuda")
    allr = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allr, t)
    for r in range(world):
        b = allr[r].cpu().numpy()
        if any(a[i].tobytes() != b[i].tobytes() for i in range(len(a)) if i not in skip):
            fails.append(f"rank {rank} and rank {r} disagree")
    nf = torch.tensor([len(fails)], device="cuda")
    dist.all_reduce(nf)
    return {"ok": int(nf.item()) == 0, "world": world, "transports": transports, "checks_per_rank": checks,
            "failures": fails[:4]}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    log2n = args.log2n
    # the suite holds 91 vectors: run the TRUE length when the host has the memory for it
    need = 91 * (8 << log2n) * 1.15
    note = None
    while log2n > 16 and need > host_mem_available_bytes() * 0.8:
        log2n -= 1
        need /= 2
        note = f"host memory allows 2^{log2n} only"
    try:
        gbs, ms, sample, steps_run, checksum = run_reference_suite(log2n, args.steps, args.warmup, threads)
    except FileNotFoundError as e:
        print(json.dumps({"impl": "reference", "unavailable": str(e)}))
        return 0
    cfg_args = argparse.Namespace(**vars(args))
    cfg_args.log2n = log2n
    line = {
        "impl": "reference", "metric": "N_Vector op suite throughput (algorithmic GB/s)", "value": round(gbs, 2),
        "unit": "GB/s", "n_gpus": args.gpus, "gpus_used": 0, "host_cores": threads, "ranks_run": 1,
        "steps": steps_run, "warmup": args.warmup,
        "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(cfg_args),
        "cpu_baseline": {"value": round(gbs, 2), "unit": "GB/s", "cores": threads, "kind": "reference",
                         "sample": sample},
        "e2e": {"value": round(gbs, 2), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "result_checksum": checksum,
        "note": note or "n_gpus is the launch shape the driver asked for; this arm runs on the host cores of rank 0 "
                        "only (gpus_used 0), on the same per-GPU workload",
    }
    print(json.dumps(line))
    return 0


def workload_config(args):
    return {
        "workload": "benchmarks/nvector performance suite (all ops incl. fused/vector-array), "
                    f"length 2^{args.log2n} per GPU, nvecs={NVECS}, nsums={NSUMS}, fused ops enabled",
        "length_per_gpu": 1 << args.log2n,
        "nvecs": NVECS, "nsums": NSUMS,
        "driver": "apps/nvector_perf (C loop over the N_Vector ops table, as benchmarks/nvector does)",
        "partition": "contiguous 1-D block per GPU (MPIPlusX pattern); reductions fold the ranks' partials over "
                     "NVLink peer memory inside the reduction kernel (NCCL allreduce as fallback); no other communication",
        "cache": f"inputs {'larger than' if args.log2n >= 24 else 'vs'} L2: every op streams {(8 << args.log2n) / 2**20:g} MiB per "
                 f"operand (126 MB L2), 91 distinct vectors ({91 * (8 << args.log2n) / 2**30:.1f} GiB) rotate through the suite",
    }


# --------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------
def bind_to_gpu_numa_node(local_rank):
    """One process per GPU: run this rank's host thread (and therefore its pinned allocations, which
    follow the first-touch / local NUMA policy) on the NUMA node the GPU's PCIe link hangs off.  With 8
    ranks uploading 1 GiB per step each, pinned buffers on the wrong socket turn the PCIe copies into
    cross-socket traffic.  Best effort: any failure leaves the affinity untouched."""
    try:
        r = subprocess.run(["nvidia-smi", f"--id={local_rank}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                           capture_output=True, text=True, timeout=20)
        bus = r.stdout.strip().lower()
        if not bus:
            return None
        if bus.count(":") == 2 and len(bus.split(":")[0]) == 8:   # nvidia-smi prints an 8-digit PCI domain
            bus = bus[4:]
        node = int(Path(f"/sys/bus/pci/devices/{bus}/numa_node").read_text())
        if node < 0:
            return {"node": node, "bound": False}
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"node": node, "bound": False}
        os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": len(cpus), "bound": True}
    except Exception as e:  # noqa: BLE001
        return {"bound": False, "error": f"{type(e).__name__}: {e}"}



def b200_vectors(P, lib, ctx, n, world, rank, nv=NVECS, ns=NSUMS, keep_host_x=True, seed=1234):
    """the suite's vectors on NVECTOR_B200 (device memory), same data as cpu_vectors(): four base
    arrays are uploaded once and every vector is scale(i) * base, formed ON the device by N_VScale;
    only the X vectors (re-uploaded by the e2e leg) keep a pinned host mirror."""
    import numpy as np

    def newvec():
        v = P.new(n, ctx, P.DEVICE, fused=True)
        if world > 1:
            assert lib.N_VMakeDistributed_B200(v, n * world) == 0
        return v

    vec = alloc_vectors(newvec, nv, ns)
    u, w, idm, cn = base_arrays(n, seed + rank)
    base = newvec()
    hb = P.host(base, n)

    def upload(arr):
        hb[...] = arr
        P.to_device(base)

    upload(u)
    xs = set(vec["X"])
    for i, v in enumerate(all_handles(vec)):
        if v in (vec["W"], vec["ID"], vec["CN"]):
            continue
        P.Scale(vec_scale(i), base, v)
        if keep_host_x and v in xs:
            np.multiply(u, vec_scale(i), out=P.host(v, n))
    upload(w)
    P.Scale(vec_scale(1), base, vec["W"])
    upload(idm)
    P.Scale(1.0, base, vec["ID"])
    upload(cn)
    P.Scale(1.0, base, vec["CN"])
    lib.b200vec_ctx_sync(ctx)
    P.Destroy(base)
    return vec


def run_sweep(P, lib, perf, ctx, world, rank, dist, torch, peak, lengths, cpu_lengths, with_cpu):
    """BASELINE configs[2]: one representative op per kernel class over vector lengths 2^16 .. 2^30
    per GPU (C driver, CUDA events, the API call incl. the host hand-off for the scalar ops), and the
    reference's nvector_serial / nvector_openmp on this box's host cores at the same lengths (bounded)."""
    classes = ["N_VLinearSum-9", "N_VScale-4", "N_VConst", "N_VDotProd", "N_VMaxNorm", "N_VWrmsNormMask",
               "N_VLinearCombination-3", "N_VScaleAddMulti-2", "N_VDotProdMulti"]
    out = {"ops": classes, "lengths": {}, "timing": "C driver loop, CUDA events; reps back to back on one operand set "
           "(> L2 from 2^22 x 3 operands on; smaller lengths are L2 / launch-latency figures)"}
    free, _ = torch.cuda.mem_get_info()
    if dist is not None:   # every rank must pick the same vector count: the reductions are collectives
        t = torch.tensor([float(free)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        free = int(t.item())
    for L in lengths:
        n = 1 << L
        nv = 8
        while nv > 2 and (5 * nv + 5) * n * 8 > free * 0.85:
            nv //= 2
        if (5 * nv + 5) * n * 8 > free * 0.85:
            out["lengths"][f"2^{L}"] = {"skipped": "does not fit HBM"}
            continue
        newvec_count = [0]

        def newvec():
            newvec_count[0] += 1
            v = P.new(n, ctx, P.DEVICE, fused=True)
            if world > 1:
                assert lib.N_VMakeDistributed_B200(v, n * world) == 0
            P.Const(0.75 + 0.01 * (newvec_count[0] % 17), v)
            return v

        X, Y, Z = [newvec() for _ in range(nv)], [newvec() for _ in range(nv)], [newvec() for _ in range(nv)]
        vec = {"X": X, "Y": Y, "Z": Z, "S": newvec(), "T": newvec(), "W": newvec(), "ID": newvec(), "CN": newvec(),
               "YY": [Y], "ZZ": [Z]}
        suite = Suite(perf, vec)
        row = {}
        for name in classes:
            i = suite.names.index(name)
            suite.op(i, 3)
            torch.cuda.synchronize()
            reps = int(max(5, min(200, 40e9 / (suite.bpe[i] * n))))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if dist is not None:
                dist.barrier()
            e0.record()
            suite.op(i, reps)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / reps * 1e3
            if dist is not None:
                t = torch.tensor([us], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                us = float(t.item())
            gbs = suite.bpe[i] * n / us / 1e3
            row[name] = {"us": round(us, 2), "GBs": round(gbs, 1), "frac": round(gbs / peak, 3)}
        suite.check()
        out["lengths"][f"2^{L}"] = {"nvecs": nv, "ops": row}
        for v in X + Y + Z + [vec["S"], vec["T"], vec["W"], vec["ID"], vec["CN"]]:
            P.Destroy(v)
        lib.b200vec_ctx_sync(ctx)
    if with_cpu and rank == 0:
        try:
            ref = load_reference()
            sctx = C.c_void_p()
            assert ref.SUNContext_Create(0, C.byref(sctx)) == 0
            threads = os.cpu_count() or 1
            cpu = {"cores": threads, "lengths": {}}
            for L in cpu_lengths:
                n = 1 << L
                if 45 * n * 8 * 1.2 > host_mem_available_bytes() * 0.7:
                    cpu["lengths"][f"2^{L}"] = {"skipped": "host memory"}
                    continue
                row = {}
                for kind, th in (("serial", 1), ("openmp", threads)):
                    import numpy as np

                    def newvec():
                        v = ref.N_VNew_OpenMP(n, th, sctx) if th > 1 else ref.N_VNew_Serial(n, sctx)
                        (ref.N_VEnableFusedOps_OpenMP if th > 1 else ref.N_VEnableFusedOps_Serial)(v, 1)
                        np.ctypeslib.as_array(ref.N_VGetArrayPointer(v), shape=(n,))[...] = 0.75
                        return v

                    X, Y, Z = [newvec() for _ in range(8)], [newvec() for _ in range(8)], [newvec() for _ in range(8)]
                    vec = {"X": X, "Y": Y, "Z": Z, "S": newvec(), "T": newvec(), "W": newvec(), "ID": newvec(),
                           "CN": newvec(), "YY": [Y], "ZZ": [Z]}
                    suite = Suite(perf, vec)
                    for name in classes:
                        i = suite.names.index(name)
                        suite.op(i, 1)
                        reps = int(max(2, min(200, 1.5e9 / (suite.bpe[i] * n))))
                        t0 = time.perf_counter()
                        suite.op(i, reps)
                        us = (time.perf_counter() - t0) / reps * 1e6
                        row.setdefault(name, {})[kind] = {"us": round(us, 1), "GBs": round(suite.bpe[i] * n / us / 1e3, 2)}
                    for v in X + Y + Z + [vec["S"], vec["T"], vec["W"], vec["ID"], vec["CN"]]:
                        ref.N_VDestroy(v)
                cpu["lengths"][f"2^{L}"] = row
            out["cpu"] = cpu
        except Exception as e:  # reported, never required
            out["cpu"] = {"unavailable": f"{type(e).__name__}: {e}"}
    # summary: how many of the 9 classes reach 0.8 of the measured peak, and the scalar-op latency floor
    summ = {}
    for key, val in out["lengths"].items():
        if "ops" in val:
            summ[key] = sum(1 for o in val["ops"].values() if o["frac"] >= 0.8)
    out["classes_ge_0.8_of_peak"] = summ
    return out


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
    numa = bind_to_gpu_numa_node(local_rank) if not args.no_numa_bind else None
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
    perf = load_perf()
    n = 1 << args.log2n
    for name, res, argt in (("N_VCopyToDeviceAsync_B200", None, [V]), ("N_VCopyJoin_B200", None, [V])):
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, argt

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

    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"

    # ---- distributed exact-answer parity on both transports, before anything is timed
    parity = None
    if world > 1:
        try:
            parity = dist_parity(P, lib, ctx, rank, world, dist, torch)
        except Exception as e:
            parity = {"ok": False, "world": world, "error": f"{type(e).__name__}: {e}"}

    vec = b200_vectors(P, lib, ctx, n, world, rank)
    suite = Suite(perf, vec)
    bytes_per_step = suite.bytes_per_elt_step * n

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        """fn() bracketed by barrier+sync, CUDA events on the launching (legacy) stream, max over ranks."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- value: operands resident in HBM; K steps in ONE call of the C driver
    suite.step(args.warmup)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lib.b200vec_ctx_set_tuning(ctx, b"count_launches", 1)
    ms_step = timed(lambda: suite.step(args.steps)) / args.steps
    launches = int(lib.b200vec_ctx_launch_count(ctx))
    lib.b200vec_ctx_set_tuning(ctx, b"count_launches", 0)
    suite.check()
    result_value = suite.result("N_VWrmsNorm(result)")

    # ---- e2e: every step uploads the nvecs input vectors from pinned host memory and reads a result
    # vector + the scalars back.  Double-buffered: two device copies of X; the upload of step k+1 runs on
    # the context's copy stream while the kernels of step k run; the first upload is not overlapped and
    # is inside the timed region, so K uploads + K steps + K read-backs are timed.
    h2d = NVECS * n * 8
    d2h = n * 8 + 8 * 24
    XB = []
    for v in vec["X"]:
        c = P.Clone(v)
        lib.N_VSetHostArrayPointer_B200(P.host(v, n).ctypes.data_as(C.POINTER(C.c_double)), c)  # shared pinned mirror
        XB.append(c)
    suites = [suite, Suite(perf, vec, X=XB)]
    xsets = [vec["X"], XB]

    def upload(which):
        for v in xsets[which]:
            lib.N_VCopyToDeviceAsync_B200(v)

    def run_e2e(k):
        upload(0)
        lib.N_VCopyJoin_B200(vec["X"][0])
        for i in range(k):
            cur = i & 1
            if i + 1 < k:
                upload(1 - cur)
            suites[cur].step(1)
            P.from_device(vec["Z"][1])          # D2H of the step's result vector (+ the scalars already on the host)
            lib.N_VCopyJoin_B200(vec["X"][0])    # the next step's inputs must have landed (stream-side wait)

    run_e2e(min(3, args.warmup))
    ms_e2e = timed(lambda: run_e2e(args.steps)) / args.steps
    suites[1].check()
    clocks = sampler.stop() if rank == 0 else None

    value = world * bytes_per_step / (ms_step * 1e-3) / 1e9
    e2e_value = world * bytes_per_step / (ms_e2e * 1e-3) / 1e9

    # ---- per-op timings: each op alone in a C loop, CUDA events ("API": for a scalar-returning op the
    # call blocks until the host has the value).  EVERY rank runs this loop -- the reducing ops of a
    # distributed vector are collectives (SPMD) -- rank 0 reports
    per_op = {}
    reps = 10
    for i, name in enumerate(suite.names):
        suite.op(i, 1)
        barrier()   # the streaming ops before a collective op let the ranks drift apart: start each op aligned
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        suite.op(i, reps)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        per_op[name] = {"us": round(us, 2), "GBs": round(suite.bpe[i] * n / us / 1e3, 1),
                        "frac_of_peak": round(suite.bpe[i] * n / us / 1e3 / peak, 3)}

    # scalar-returning ops: where the API time goes.  dev_us = %globaltimer from the kernel's first CTA
    # to the publication of the result (device side, one stamped call); the rest of the API time is the
    # launch path + the PCIe hand-off + the host poll, which no kernel design removes.
    lib.b200vec_ctx_set_tuning(ctx, b"profile", 1)
    for i, name in enumerate(suite.names):
        if not suite.scalar[i]:
            continue
        devs = []
        for _ in range(3):
            lib.b200vec_ctx_set_tuning(ctx, b"prof_stamp_reset", 1)
            suite.op(i, 1)
            t0 = lib.b200vec_ctx_get_tuning(ctx, b"prof_counter_6")
            t1 = lib.b200vec_ctx_get_tuning(ctx, b"prof_counter_7")
            if 0 < t0 < t1:
                devs.append((t1 - t0) * 1e-3)
        if devs:
            d = sorted(devs)[len(devs) // 2]
            per_op[name].update({"dev_us": round(d, 2), "dev_frac_of_peak": round(suite.bpe[i] * n / d / 1e3 / peak, 3),
                                 "launch_and_return_us": round(per_op[name]["us"] - d, 2)})
    lib.b200vec_ctx_set_tuning(ctx, b"profile", 0)

    # dominant kernel by share of the step (profiles/r01_launches_summary.md: 28 %):
    # k_scaleadd_rows<4>, launched 3x per step (N_VScaleAddMulti-1/-2, 136 B/elt, and
    # N_VScaleAddMultiVectorArray, 576 B/elt).  achieved = algorithmic bytes of
    # those launches / their CUDA-event time, measured live above.
    dom_ops = ["N_VScaleAddMulti-1", "N_VScaleAddMulti-2", "N_VScaleAddMultiVectorArray"]
    dom_bytes = {name: suite.bpe[suite.names.index(name)] * n for name in dom_ops}
    dom_us = sum(per_op[k]["us"] for k in dom_ops)
    dom_gbs = sum(dom_bytes.values()) / dom_us / 1e3
    roofline = {"bound": "hbm", "kernel": "k_scaleadd_rows<4> (N_VScaleAddMulti, N_VScaleAddMultiVectorArray)",
                "achieved": round(dom_gbs, 1), "peak": peak, "unit": "GB/s", "frac": round(dom_gbs / peak, 4),
                "traffic": None, "peak_source": peak_src,
                "launches_per_step": 3,
                "algorithmic_bytes_per_launch": dom_bytes,
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

    # the suite's vectors are no longer needed: the legs below allocate their own
    for v in XB:
        P.Destroy(v)
    for v in all_handles(vec):
        P.Destroy(v)
    lib.b200vec_ctx_sync(ctx)

    def leg_clock_min(d):
        try:
            return min(g["sm_mhz_min"] for g in d["clocks"].values())
        except Exception:
            return None

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
        # the config's other nonlinear solver: Anderson-accelerated fixed point (no linear solver)
        try:
            # (the stiff reaction terms hold a fixed-point iteration to ~1e-5 steps: 4246 steps for tf = 0.05,
            #  so this run is bounded much shorter than the Newton one; compare ms_per_step)
            fp = run_ar3d(ctx, world, args, nls="fixedpoint", tf=args.ar3d_fp_tf)
            ar3d["fixedpoint"] = {k: fp[k] for k in ("workload", "solve_s", "steps", "step_attempts", "fe_evals", "fi_evals",
                                                       "nls_iters", "ms_per_step", "urms", "vrms", "wrms")}
        except Exception as e:
            ar3d["fixedpoint"] = {"unavailable": f"{type(e).__name__}: {e}"}

    # ---- Krylov Gram-Schmidt built on the ops: the reference's unmodified
    # SUNClassicalGS / SUNModifiedGS, and the fused SUNClassicalGS_B200, on a basis of this vector; collective
    gs = None
    if not args.no_gs:
        try:
            sys.path.insert(0, str(ROOT / "tools"))
            import gs_bench

            gs = gs_bench.run(args.log2n, maxl=5, reps=5, cpu_log2n=args.cpu_log2n, with_ref_cuda=False,
                              with_cpu=(world == 1 and not args.no_cpu_baseline), b200_ctx=ctx, rank=rank, world=world)
            for k in gs["b200"]:
                gs["b200"][k]["cycle_frac_of_peak"] = round(gs["b200"][k]["cycle_GBs"] / peak, 3)
        except Exception as e:
            gs = {"unavailable": f"{type(e).__name__}: {e}"}

    # ---- CVODE's fused-kernel plugin (SURVEY f-N3): libsundials_cvode_fused_b200.so beside the unfused path (the
    # reference's stubs library driving this vector through the ops table), N = 1 only
    cvf = None
    if world == 1 and not args.no_cvfused:
        try:
            sys.path.insert(0, str(ROOT / "tools"))
            import cvfused_bench

            cvf = cvfused_bench.run(args.log2n, reps=10, with_ref_cuda=False)
        except Exception as e:
            cvf = {"unavailable": f"{type(e).__name__}: {e}"}

    # ---- length sweep (BASELINE config 3), bounded in the default run; collective
    sweep = None
    if not args.no_sweep:
        try:
            lengths = list(range(16, 31, 2)) if args.sweep else [16, 20, 24, 28]
            cpu_lengths = [L for L in lengths if L <= 26] if args.sweep else [16, 20, 24]
            sweep = run_sweep(P, lib, perf, ctx, world, rank, dist, torch, peak, lengths, cpu_lengths,
                              with_cpu=(not args.no_cpu_baseline and (world == 1 or args.sweep)))
        except Exception as e:
            sweep = {"unavailable": f"{type(e).__name__}: {e}"}

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- CVODE cvDiurnal_kry (BASELINE config 2), single GPU / host program
    cvd = None
    if not args.no_cvdiurnal:
        try:
            cvd = run_cvdiurnal()
        except Exception as e:
            cvd = {"unavailable": f"{type(e).__name__}: {e}"}

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
            g_omp, _, sample_omp, _, chk = run_reference_suite(args.cpu_log2n, 2, 1, threads, budget_s=25.0)
            g_ser, _, sample_ser, _, _ = run_reference_suite(min(args.cpu_log2n, 20), 1, 0, 1, budget_s=15.0)
            cpu = {"value": round(g_omp, 2), "unit": "GB/s", "cores": threads, "kind": "reference",
                   "sample": sample_omp + " (bounded sample of the workload)", "serial_1core_GBs": round(g_ser, 2),
                   "serial_sample": sample_ser}
            # the same step on the same data at the sample's length on the GPU: the suite's result checksum
            # (WRMS norm of the last output vector) must agree with the reference's own vector
            v2 = b200_vectors(P, lib, ctx, 1 << args.cpu_log2n, 1, 0, keep_host_x=False)
            s2 = Suite(perf, v2)
            s2.step(1)
            s2.check()
            mine = s2.result("N_VWrmsNorm(result)")
            for v in all_handles(v2):
                P.Destroy(v)
            cpu["checksum_parity"] = {"length": f"2^{args.cpu_log2n}", "b200": mine, "reference_nvector_openmp": chk,
                                      "rel_diff": abs(mine - chk) / abs(chk) if chk else None,
                                      "ok": bool(chk) and abs(mine - chk) <= 1e-13 * abs(chk)}
        except Exception as e:  # the baseline is reported, never required
            cpu = {"value": None, "unit": "GB/s", "cores": os.cpu_count(), "kind": "reference",
                   "sample": f"unavailable: {e}"}

    # ---- compact record of the legs, printed LAST so that it survives any tail cut of the line
    legs = {}
    if parity is not None:
        legs["dist_parity"] = {"ok": parity.get("ok"), "world": world, "transports": parity.get("transports")}
    if diffusion and "solve_s" in diffusion:
        legs["diffusion_2D"] = {"ms_per_ls_iter": diffusion["ms_per_ls_iter"], "ls_iters": diffusion["ls_iters"],
                                "solve_s": diffusion["solve_s"], "sm_mhz_min": leg_clock_min(diffusion)}
    if ar3d and "solve_s" in ar3d:
        legs["advection_reaction_3D"] = {"solve_s": ar3d["solve_s"], "steps": ar3d["steps"], "fe": ar3d["fe_evals"],
                                         "nni": ar3d["nls_iters"], "nli": ar3d["ls_iters"],
                                         "sm_mhz_min": leg_clock_min(ar3d)}
        fp = ar3d.get("fixedpoint") or {}
        if "solve_s" in fp:
            legs["advection_reaction_3D"]["fixedpoint"] = {"ms_per_step": fp["ms_per_step"], "steps": fp["steps"],
                                                           "nni": fp["nls_iters"]}
    if gs and "b200" in gs:
        legs["gram_schmidt"] = {k: gs["b200"][k]["cycle_frac_of_peak"] for k in gs["b200"]}
    if cvf is not None and "functions" in cvf:
        short = {"cvEwtSetSS_fused": "ewtSS", "cvEwtSetSV_fused": "ewtSV", "cvCheckConstraints_fused": "constr",
                 "cvNlsResid_fused": "nlsres", "cvDiagSetup_formY": "formY", "cvDiagSetup_buildM": "buildM",
                 "cvDiagSolve_updateM": "updM"}
        # [fraction of the HBM peak, speed-up over the unfused op sequence on the same vector]
        legs["cvode_fused"] = {short[k]: [v["fused_frac_of_peak"], v["speedup_vs_unfused"]] for k, v in cvf["functions"].items()}
    if cvd and "b200_pinned_s" in cvd:
        legs["cvDiurnal_kry"] = {"b200_s": cvd["b200_pinned_s"], "serial_s": cvd.get("serial_1core_s"),
                                 "identical": cvd.get("stdout_identical_to_serial_golden")}
    if sweep and "lengths" in sweep:
        sw = {"ge0.8_of_9": sweep.get("classes_ge_0.8_of_peak")}
        for L in ("2^16", "2^20"):
            try:
                sw[f"dot_us_{L}"] = sweep["lengths"][L]["ops"]["N_VDotProd"]["us"]
            except KeyError:
                pass
        legs["sweep"] = sw
    red = ["N_VDotProd", "N_VMaxNorm", "N_VMin", "N_VL1Norm", "N_VWrmsNorm", "N_VWrmsNormMask", "N_VInvTest",
           "N_VConstrMask", "N_VMinQuotient"]
    if cpu and "checksum_parity" in cpu:
        legs["checksum_vs_reference_ok"] = cpu["checksum_parity"]["ok"]
    legs["reductions_api_frac"] = {k[3:]: per_op[k]["frac_of_peak"] for k in red}
    legs["reductions_dev_frac"] = {k[3:]: per_op[k].get("dev_frac_of_peak") for k in red}

    lib.b200vec_comm_transport.restype = C.c_char_p
    line = {
        "metric": "N_Vector op suite throughput (algorithmic GB/s)", "value": round(value, 1), "unit": "GB/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "e2e": {"value": round(e2e_value, 1), "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": round(ms_e2e, 4),
                "overlap": "uploads double-buffered on the context's copy stream (N_VCopyToDeviceAsync_B200)"},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "clocks": clocks,
        "reduction_transport": lib.b200vec_comm_transport(ctx).decode(),
        "numa": numa,
        "result_checksum": result_value,
        "ops_per_step": suite.nops,
        "dist_parity": parity,
        "diffusion_2D": diffusion,
        "advection_reaction_3D": ar3d,
        "gram_schmidt": gs,
        "cvDiurnal_kry": cvd,
        "cvode_fused": cvf,
        "sweep": sweep,
        "per_op": per_op,
        "legs": legs,
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
    ap.add_argument("--ar3d-fp-tf", type=float, default=5e-4, help="final time of the fixed-point variant of the leg")
    ap.add_argument("--no-gs", action="store_true", help="skip the Gram-Schmidt leg")
    ap.add_argument("--no-cvfused", action="store_true", help="skip the CVODE fused-kernel plugin leg")
    ap.add_argument("--no-sweep", action="store_true", help="skip the (bounded) length sweep leg")
    ap.add_argument("--sweep", action="store_true", help="full length sweep 2^16 .. 2^30 with CPU columns up to 2^26")
    ap.add_argument("--no-cvdiurnal", action="store_true", help="skip the CVODE cvDiurnal_kry leg")
    ap.add_argument("--only-suite", action="store_true", help="op suite only: skip every leg")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not bind the rank to its GPU's NUMA node")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.only_suite:
        args.no_diffusion = args.no_ar3d = args.no_gs = args.no_sweep = args.no_cvdiurnal = args.no_cvfused = True
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
