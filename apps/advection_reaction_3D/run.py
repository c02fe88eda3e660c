#!/usr/bin/env python
"""Driver of the re-hosted advection_reaction_3D benchmark (apps/advection_reaction_3D/ar3d_b200.cu).

Single GPU:   python apps/advection_reaction_3D/run.py --npts 256 --method ARK-IMEX --nls newton
N GPUs:       python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
                  --master-port 29571 apps/advection_reaction_3D/run.py --npts 512 --method ARK-IMEX

Options carry the reference's names (benchmarks/advection_reaction_3D/raja/README.md,
advection_reaction_3D.cpp:283-437).  --npts is the GLOBAL mesh size per direction; ranks own
slabs in x (the reference's --npxyz N 1 1).  torch.distributed is used only to hand the NCCL
unique id of the vector's communicator to the other ranks.  Unlike the reference the solution
files are written only with --save (the reference needs --dont-save to skip them).
stdout: the reference's screen output (rank 0); with --json one JSON line of statistics follows.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
LIB = Path(__file__).resolve().parent / "_build" / "libar3d_b200.so"

METHODS = {"ERK": 0, "ARK-DIRK": 1, "ARK-IMEX": 2, "CV-BDF": 3, "CV-ADAMS": 4}
NLS = {"newton": 0, "tl-newton": 1, "fixedpoint": 2}
RHS_ADVECTION, RHS_REACTION, RHS_ADVECTION_REACTION = 0, 1, 2


class Opts(C.Structure):
    _fields_ = [("npts", C.c_int64), ("xmax", C.c_double), ("A", C.c_double), ("B", C.c_double),
                ("k1", C.c_double), ("k2", C.c_double), ("k3", C.c_double), ("k4", C.c_double),
                ("k5", C.c_double), ("k6", C.c_double), ("c", C.c_double), ("method", C.c_int), ("nls", C.c_int),
                ("order", C.c_int), ("fpaccel", C.c_int), ("precond", C.c_int), ("fused", C.c_int),
                ("t0", C.c_double), ("tf", C.c_double), ("rtol", C.c_double), ("atol", C.c_double),
                ("nout", C.c_int), ("save", C.c_int), ("outputdir", C.c_char * 1024), ("output", C.c_int),
                ("force_generic", C.c_int), ("planes_per_cta", C.c_int), ("fused_ewt", C.c_int)]


class Stats(C.Structure):
    _fields_ = [(k, C.c_long) for k in ("nst", "nst_a", "netf", "nfe", "nfi", "nni", "ncnf", "nli", "npsol",
                                         "nnlfi")] + \
               [(k, C.c_double) for k in ("t_final", "urms", "vrms", "wrms", "evolve_seconds", "setup_seconds",
                                          "rhs_seconds")] + \
               [("rhs_calls", C.c_long), ("psolve_calls", C.c_long), ("neq", C.c_int64), ("neq_loc", C.c_int64),
                ("nranks", C.c_int)]


_lib_handle = None


def load():
    global _lib_handle
    if _lib_handle is not None:
        return _lib_handle
    from sundials_b200 import _lib

    _lib.load()  # libsundials_nvecb200.so first (RTLD_GLOBAL), then the app
    if not LIB.exists():
        raise FileNotFoundError(f"{LIB} missing: run `make -C apps/advection_reaction_3D` (needs nvcc + SUNDIALS headers)")
    lib = C.CDLL(str(LIB), mode=C.RTLD_GLOBAL)
    lib.b200_ar3d_default_opts.argtypes = [C.POINTER(Opts)]
    lib.b200_ar3d_run.restype = C.c_int
    lib.b200_ar3d_run.argtypes = [C.c_void_p, C.POINTER(Opts), C.POINTER(Stats)]
    lib.b200_ar3d_plan_create.restype = C.c_int
    lib.b200_ar3d_plan_create.argtypes = [C.c_void_p, C.POINTER(Opts), C.POINTER(C.c_void_p)]
    lib.b200_ar3d_plan_destroy.argtypes = [C.c_void_p]
    lib.b200_ar3d_plan_local_neq.restype = C.c_int64
    lib.b200_ar3d_plan_local_neq.argtypes = [C.c_void_p]
    lib.b200_ar3d_plan_is_fast.restype = C.c_int
    lib.b200_ar3d_plan_is_fast.argtypes = [C.c_void_p]
    lib.b200_ar3d_set_ic.restype = C.c_int
    lib.b200_ar3d_set_ic.argtypes = [C.c_void_p, C.c_void_p]
    lib.b200_ar3d_component_mask.restype = C.c_int
    lib.b200_ar3d_component_mask.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.b200_ar3d_rhs.restype = C.c_int
    lib.b200_ar3d_rhs.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.b200_ar3d_psolve.restype = C.c_int
    lib.b200_ar3d_psolve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
    _lib_handle = lib
    return lib


def make_opts(**kw) -> Opts:
    lib = load()
    o = Opts()
    lib.b200_ar3d_default_opts(C.byref(o))
    for k, v in kw.items():
        if k == "outputdir":
            o.outputdir = str(v).encode()
        elif k == "method" and isinstance(v, str):
            o.method = METHODS[v]
        elif k == "nls" and isinstance(v, str):
            o.nls = NLS[v]
        elif k == "k":
            o.k1 = o.k2 = o.k3 = o.k4 = v
        else:
            if not hasattr(o, k):
                raise KeyError(k)
            setattr(o, k, v)
    return o


def make_context(local_rank: int, rank: int, world: int):
    """b200vec context (+ communicator when world > 1; the unique id travels by torch.distributed)."""
    import torch

    from sundials_b200 import _lib

    lib = _lib.load()
    ctx = C.c_void_p()
    _lib.check(lib.b200vec_ctx_create(C.byref(ctx), local_rank, None), "ctx_create")
    if world > 1:
        import torch.distributed as dist

        idbuf = (C.c_ubyte * _lib.UNIQUE_ID_BYTES)()
        if rank == 0:
            _lib.check(lib.b200vec_comm_get_unique_id(idbuf), "comm_get_unique_id")
        t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, src=0)
        idbuf = (C.c_ubyte * _lib.UNIQUE_ID_BYTES).from_buffer_copy(bytes(t.cpu().tolist()))
        _lib.check(lib.b200vec_comm_init(ctx, idbuf, rank, world), "comm_init")
    return ctx


def run(ctx, **kw) -> dict:
    lib = load()
    o = make_opts(**kw)
    st = Stats()
    sys.stdout.flush()
    rc = lib.b200_ar3d_run(ctx, C.byref(o), C.byref(st))
    if rc != 0:
        raise RuntimeError(f"b200_ar3d_run failed ({rc})")
    return {k: getattr(st, k) for k, _ in Stats._fields_}


class Plan:
    """The building blocks on torch device tensors (tests, profiling)."""

    def __init__(self, ctx, **kw):
        self.lib = load()
        self.opts = make_opts(**kw)
        self.h = C.c_void_p()
        if self.lib.b200_ar3d_plan_create(ctx, C.byref(self.opts), C.byref(self.h)) != 0:
            raise RuntimeError("b200_ar3d_plan_create failed")
        self.neq_loc = self.lib.b200_ar3d_plan_local_neq(self.h)
        self.fast = bool(self.lib.b200_ar3d_plan_is_fast(self.h))

    def close(self):
        if self.h:
            self.lib.b200_ar3d_plan_destroy(self.h)
            self.h = None

    def _chk(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc})")

    def set_ic(self, y):
        self._chk(self.lib.b200_ar3d_set_ic(self.h, y.data_ptr()), "set_ic")

    def component_mask(self, comp, m):
        self._chk(self.lib.b200_ar3d_component_mask(self.h, comp, m.data_ptr()), "component_mask")

    def rhs(self, which, y, f):
        self._chk(self.lib.b200_ar3d_rhs(self.h, which, y.data_ptr(), f.data_ptr()), "rhs")

    def psolve(self, y, b, x, gamma):
        self._chk(self.lib.b200_ar3d_psolve(self.h, y.data_ptr(), b.data_ptr(), x.data_ptr(), gamma), "psolve")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--npts", type=int, default=100)
    ap.add_argument("--xmax", type=float, default=1.0)
    ap.add_argument("--A", type=float, default=1.0)
    ap.add_argument("--B", type=float, default=3.5)
    ap.add_argument("--k", type=float, default=1.0)
    ap.add_argument("--c", type=float, default=0.01)
    ap.add_argument("--method", default="ARK-DIRK", choices=list(METHODS))
    ap.add_argument("--nls", default="newton", choices=list(NLS))
    ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--fpaccel", type=int, default=3)
    ap.add_argument("--nopre", action="store_true")
    ap.add_argument("--fused", action="store_true")
    ap.add_argument("--tf", type=float, default=10.0)
    ap.add_argument("--rtol", type=float, default=1e-6)
    ap.add_argument("--atol", type=float, default=1e-9)
    ap.add_argument("--nout", type=int, default=10)
    ap.add_argument("--save", action="store_true", help="write u/v/w.<rank>.txt, t.000000.txt, mesh.txt")
    ap.add_argument("--dont-save", action="store_true", help="accepted for compatibility (the default here)")
    ap.add_argument("--output-dir", default=".")
    ap.add_argument("--quiet", action="store_true")
    ap.add_argument("--generic", action="store_true", help="force the one-node-per-thread RHS kernel")
    ap.add_argument("--planes-per-cta", type=int, default=0)
    ap.add_argument("--exact-threshold", type=int, default=None,
                    help="vector length up to which reductions sum in serial order (<= 4096)")
    ap.add_argument("--json", action="store_true")
    ap.add_argument("--profile", action="store_true",
                    help="device-side wait times of the cross-rank exchanges (reductions, halo, acknowledges), per rank")
    ap.add_argument("--p2p", type=int, default=None, help="0: ncclAllReduce after the reduction kernel instead of peer memory")
    a = ap.parse_args()

    import torch

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    lrank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("advection_reaction_3D needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(lrank)
    if world > 1:
        import datetime

        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lrank}"),
                                timeout=datetime.timedelta(seconds=180))
    ctx = make_context(lrank, rank, world)
    if a.exact_threshold is not None:
        from sundials_b200 import _lib

        _lib.check(_lib.load().b200vec_ctx_set_tuning(ctx, b"exact_threshold", a.exact_threshold), "set_tuning")
    if a.profile or a.p2p is not None:
        from sundials_b200 import _lib

        L = _lib.load()
        if a.p2p is not None:
            _lib.check(L.b200vec_ctx_set_tuning(ctx, b"p2p", a.p2p), "set_tuning(p2p)")
        if a.profile:
            _lib.check(L.b200vec_ctx_set_tuning(ctx, b"profile", 1), "set_tuning(profile)")
    st = run(ctx, npts=a.npts, xmax=a.xmax, A=a.A, B=a.B, k=a.k, c=a.c, method=a.method, nls=a.nls, order=a.order,
             fpaccel=a.fpaccel, precond=0 if a.nopre else 1, fused=1 if a.fused else 0, tf=a.tf, rtol=a.rtol,
             atol=a.atol, nout=a.nout, save=1 if a.save else 0, outputdir=a.output_dir, output=0 if a.quiet else 1,
             force_generic=1 if a.generic else 0, planes_per_cta=a.planes_per_cta)
    if a.profile:
        g = lambda i: L.b200vec_ctx_get_tuning(ctx, f"prof_counter_{i}".encode())  # noqa: E731
        print("PROFILE " + json.dumps({
            "rank": rank, "nranks": world, "evolve_s": round(st["evolve_seconds"], 4), "steps": st["nst"],
            "reduction_exchange_ms": round(g(0) / 1e6, 3), "reduction_exchanges": g(1),
            "halo_wait_ms_summed_over_ctas": round(g(2) / 1e6, 3), "halo_waiting_ctas": g(3),
            "ack_wait_ms_summed_over_ctas": round(g(4) / 1e6, 3), "ack_waiting_ctas": g(5)}), file=sys.stderr, flush=True)
    if a.json and rank == 0:
        print(json.dumps(st), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
