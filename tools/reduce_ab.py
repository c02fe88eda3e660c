"""A/B of the single-output reduction kernels (k_reduce) with and without L2 bulk prefetch
look-ahead ("l2_prefetch" tuning key = tiles of cp.async.bulk.prefetch.L2 per CTA), kernel only
(C ABI, result stays on the device), CUDA events, operands rotating over `nbuf` vectors so that
no repetition finds its input in the 126 MB L2.

    python tools/reduce_ab.py [--log2n 22,24,26] [--reps 40] > gpurun_out/reduce_ab.json
"""
import argparse
import ctypes as C
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from sundials_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", default="22,24,26")
    ap.add_argument("--reps", type=int, default=40)
    ap.add_argument("--settings", default="0,1,2")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    lib = _lib.load()
    ctx = C.c_void_p()
    _lib.check(lib.b200vec_ctx_create(C.byref(ctx), 0, None), "ctx_create")
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    peak = float(peaks.get("hbm_gbs", 6650.0))
    out = {"peak_GBs": peak, "reps": a.reps, "rows": []}
    for lg in [int(x) for x in a.log2n.split(",")]:
        n = 1 << lg
        nbuf = max(4, min(16, (1 << 31) // (8 * n)))      # >= 2 GiB of rotating operands (or 4 vectors)
        g = torch.Generator(device="cuda").manual_seed(5)
        bufs = [torch.rand(n, dtype=torch.float64, device="cuda", generator=g) + 0.5 for _ in range(nbuf)]
        outv = torch.empty(n, dtype=torch.float64, device="cuda")
        p = [b.data_ptr() for b in bufs]
        ops = {
            "max_norm(8)": (8, lambda i: lib.b200vec_max_norm(ctx, p[i % nbuf], n, None)),
            "min(8)": (8, lambda i: lib.b200vec_min(ctx, p[i % nbuf], n, None)),
            "l1_norm(8)": (8, lambda i: lib.b200vec_l1_norm(ctx, p[i % nbuf], n, None)),
            "dot_prod(16)": (16, lambda i: lib.b200vec_dot_prod(ctx, p[i % nbuf], p[(i + 1) % nbuf], n, None)),
            "wsqr_sum(16)": (16, lambda i: lib.b200vec_wsqr_sum(ctx, p[i % nbuf], p[(i + 1) % nbuf], n, None)),
            "wsqr_sum_mask(24)": (24, lambda i: lib.b200vec_wsqr_sum_mask(ctx, p[i % nbuf], p[(i + 1) % nbuf],
                                                                         p[(i + 2) % nbuf], n, None)),
            "min_quotient(16)": (16, lambda i: lib.b200vec_min_quotient(ctx, p[i % nbuf], p[(i + 1) % nbuf], n, None)),
            "inv_test(16)": (16, lambda i: lib.b200vec_inv_test(ctx, p[i % nbuf], outv.data_ptr(), n, None)),
            "constr_mask(24)": (24, lambda i: lib.b200vec_constr_mask(ctx, p[i % nbuf], p[(i + 1) % nbuf],
                                                                     outv.data_ptr(), n, None)),
        }
        for name, (bpe, fn) in ops.items():
            row = {"log2n": lg, "op": name}
            for pf in [int(x) for x in a.settings.split(",")]:
                _lib.check(lib.b200vec_ctx_set_tuning(ctx, b"l2_prefetch", pf), "set_tuning")
                for i in range(5):
                    fn(i)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(a.reps):
                    fn(i)
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) / a.reps * 1e3
                row[f"pf{pf}_us"] = round(us, 2)
                row[f"pf{pf}_frac"] = round(bpe * n / us / 1e3 / peak, 3)
            out["rows"].append(row)
            print(row, file=sys.stderr, flush=True)
        del bufs, outv
        torch.cuda.empty_cache()
    _lib.check(lib.b200vec_ctx_set_tuning(ctx, b"l2_prefetch", 0), "set_tuning")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
