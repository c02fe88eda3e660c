"""Time the diffusion_2D right-hand-side kernel alone (apps/diffusion_2D, plan API) for each
look-ahead/residency variant and rows-per-CTA setting; check that all variants write the
same bits.  Run on the B200 box:  python tools/rhs_bench.py [--nx 8192 --ny 8192]
Timing: CUDA events on the context's stream, 3 warm-ups, u/f pairs rotated over buffers
larger than L2 (each call streams 16 B/node: read u once, write f once)."""
import argparse
import ctypes as C
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "apps" / "diffusion_2D"))
import run as app  # noqa: E402
from sundials_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=8192)
    ap.add_argument("--ny", type=int, default=8192)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--variants", default="la4x2,la2x4,la3x3,la6x2,la8x1")
    ap.add_argument("--rows", default="0,32,64,128")
    ap.add_argument("--forcing", default="1,0")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    ctx = app.make_context(0, 0, 1)
    lib = app.load()
    lib.b200_diffusion2d_plan_create.restype = C.c_int
    lib.b200_diffusion2d_plan_create.argtypes = [C.c_void_p, C.POINTER(app.Opts), C.POINTER(C.c_void_p)]
    lib.b200_diffusion2d_rhs.restype = C.c_int
    lib.b200_diffusion2d_rhs.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    lib.b200_diffusion2d_plan_destroy.argtypes = [C.c_void_p]
    n = a.nx * a.ny
    nbuf = max(2, (400 << 20) // (n * 8) + 1)
    us = [torch.rand(n, dtype=torch.float64, device="cuda") for _ in range(nbuf)]
    fs = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(nbuf)]
    ref = None
    peak = 6650.0
    for forcing in [int(x) for x in a.forcing.split(",")]:
        for variant in a.variants.split(","):
            for R in [int(x) for x in a.rows.split(",")]:
                os.environ["B200_DIFFUSION_VARIANT"] = variant
                o = app.Opts()
                lib.b200_diffusion2d_default_opts(C.byref(o))
                o.nx, o.ny, o.forcing, o.rows_per_cta = a.nx, a.ny, forcing, R
                plan = C.c_void_p()
                assert lib.b200_diffusion2d_plan_create(ctx, C.byref(o), C.byref(plan)) == 0
                for i in range(3):
                    lib.b200_diffusion2d_rhs(plan, 0.3, us[i % nbuf].data_ptr(), fs[i % nbuf].data_ptr())
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(a.reps):
                    lib.b200_diffusion2d_rhs(plan, 0.3, us[i % nbuf].data_ptr(), fs[i % nbuf].data_ptr())
                e1.record()
                torch.cuda.synchronize()
                t = e0.elapsed_time(e1) / a.reps * 1e3
                lib.b200_diffusion2d_rhs(plan, 0.3, us[0].data_ptr(), fs[0].data_ptr())
                torch.cuda.synchronize()
                out = fs[0].clone()
                if forcing == 1 and ref is None:
                    ref = out
                same = bool(torch.equal(out.view(torch.int64), ref.view(torch.int64))) if forcing == 1 else None
                gbs = 16.0 * n / t / 1e3
                print(f"forcing={forcing} {variant} R={R:4d}: {t:8.2f} us  {gbs:7.1f} GB/s  {gbs/peak:.3f} of fallback peak"
                      f"  bits_equal={same}", flush=True)
                lib.b200_diffusion2d_plan_destroy(plan)


if __name__ == "__main__":
    main()
