#!/usr/bin/env python
"""Driver of the re-hosted diffusion_2D benchmark (apps/diffusion_2D/diffusion2d_b200.cu).

Single GPU:   python apps/diffusion_2D/run.py --nx 8192 --ny 8192
N GPUs:       python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
                  --master-port 29551 apps/diffusion_2D/run.py --nx 8192 --ny $((8192*N))

Options carry the reference's names (benchmarks/diffusion_2D/README.md).  The mesh
(--nx, --ny) is GLOBAL; ranks own strips in y.  torch.distributed is used only to
hand the NCCL unique id of the vector's communicator to the other ranks.
stdout: the reference's table + ARKodePrintAllStats (rank 0); with --json one JSON
line of statistics and timings follows.
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
LIB = Path(__file__).resolve().parent / "_build" / "libdiffusion2d_b200.so"


class Opts(C.Structure):
    _fields_ = [("nx", C.c_int64), ("ny", C.c_int64), ("xu", C.c_double), ("yu", C.c_double), ("kx", C.c_double),
                ("ky", C.c_double), ("tf", C.c_double), ("forcing", C.c_int), ("rtol", C.c_double),
                ("atol", C.c_double), ("order", C.c_int), ("linear", C.c_int), ("ls_gmres", C.c_int),
                ("preconditioning", C.c_int), ("liniters", C.c_int), ("msbp", C.c_int), ("epslin", C.c_double),
                ("maxsteps", C.c_int), ("controller", C.c_char * 16), ("output", C.c_int), ("nout", C.c_int),
                ("fused_ops", C.c_int), ("rows_per_cta", C.c_int), ("fused_ewt", C.c_int)]


class Stats(C.Structure):
    _fields_ = [(k, C.c_long) for k in ("nst", "nst_a", "netf", "nfe", "nfi", "nni", "ncfn", "nsetups", "nli", "nlcf",
                                         "npe", "nps", "njv", "nfeLS")] + \
               [(k, C.c_double) for k in ("t_final", "urms", "max_err", "evolve_seconds", "setup_seconds",
                                          "rhs_seconds")] + \
               [("rhs_calls", C.c_long), ("nodes", C.c_int64), ("nodes_loc", C.c_int64), ("nranks", C.c_int)]


def load():
    from sundials_b200 import _lib

    _lib.load()  # libsundials_nvecb200.so first (RTLD_GLOBAL), then the app
    if not LIB.exists():
        raise FileNotFoundError(f"{LIB} missing: run `make -C apps/diffusion_2D` (needs nvcc + SUNDIALS headers)")
    lib = C.CDLL(str(LIB), mode=C.RTLD_GLOBAL)
    lib.b200_diffusion2d_default_opts.argtypes = [C.POINTER(Opts)]
    lib.b200_diffusion2d_run.restype = C.c_int
    lib.b200_diffusion2d_run.argtypes = [C.c_void_p, C.POINTER(Opts), C.POINTER(Stats)]
    return lib


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
    o = Opts()
    lib.b200_diffusion2d_default_opts(C.byref(o))
    for k, v in kw.items():
        if k == "controller":
            o.controller = v.encode()
        else:
            setattr(o, k, v)
    st = Stats()
    sys.stdout.flush()
    rc = lib.b200_diffusion2d_run(ctx, C.byref(o), C.byref(st))
    if rc != 0:
        raise RuntimeError(f"b200_diffusion2d_run failed ({rc})")
    return {k: getattr(st, k) for k, _ in Stats._fields_}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=32)
    ap.add_argument("--ny", type=int, default=32)
    ap.add_argument("--xu", type=float, default=1.0)
    ap.add_argument("--yu", type=float, default=1.0)
    ap.add_argument("--kx", type=float, default=1.0)
    ap.add_argument("--ky", type=float, default=1.0)
    ap.add_argument("--tf", type=float, default=1.0)
    ap.add_argument("--noforcing", action="store_true")
    ap.add_argument("--rtol", type=float, default=1e-5)
    ap.add_argument("--atol", type=float, default=1e-10)
    ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--controller", default="I")
    ap.add_argument("--nonlinear", action="store_true")
    ap.add_argument("--ls", default="cg", choices=["cg", "gmres"])
    ap.add_argument("--noprec", action="store_true")
    ap.add_argument("--liniters", type=int, default=20)
    ap.add_argument("--epslin", type=float, default=0.0)
    ap.add_argument("--msbp", type=int, default=0)
    ap.add_argument("--maxsteps", type=int, default=0)
    ap.add_argument("--output", type=int, default=1)
    ap.add_argument("--nout", type=int, default=20)
    ap.add_argument("--nofused", action="store_true", help="leave the fused N_Vector ops disabled")
    ap.add_argument("--rows-per-cta", type=int, default=0, help="0 = one wave of equal row blocks")
    ap.add_argument("--exact-threshold", type=int, default=None,
                    help="vector length up to which reductions sum in serial order (<= 4096)")
    ap.add_argument("--json", action="store_true")
    a = ap.parse_args()

    import torch

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    lrank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("diffusion_2D needs a CUDA device: there is no CPU fallback")
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
    st = run(ctx, nx=a.nx, ny=a.ny, xu=a.xu, yu=a.yu, kx=a.kx, ky=a.ky, tf=a.tf, forcing=0 if a.noforcing else 1,
             rtol=a.rtol, atol=a.atol, order=a.order, controller=a.controller, linear=0 if a.nonlinear else 1,
             ls_gmres=1 if a.ls == "gmres" else 0, preconditioning=0 if a.noprec else 1, liniters=a.liniters,
             epslin=a.epslin, msbp=a.msbp, maxsteps=a.maxsteps, output=a.output, nout=a.nout,
             fused_ops=0 if a.nofused else 1, rows_per_cta=a.rows_per_cta)
    if a.json and rank == 0:
        print(json.dumps(st), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
