"""diffusion_2D on ONE B200: the reference's own CUDA benchmark (mpi_gpu backend + nvector_cuda
+ nvector_mpiplusx, unmodified, recompiled for sm_100a as a 1-rank program,
oracle/_ref/bin/arkode_diffusion_2D_refcuda) against the re-host on NVECTOR_B200
(apps/diffusion_2D).  Same mesh and options.  The reference prints no timings, so both are
timed as whole-process wall clock at two final times; the difference isolates the solve:
   solve_s = wall(tf_long) - wall(tf_short),  per LS iteration = solve_s / (nli_long - nli_short).

    python tools/d2d_compare.py [--n 8192 --tf-short 1e-6 --tf-long 1e-4] > gpurun_out/d2d_compare.json
"""
import argparse
import json
import re
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def stat(text, name):
    m = re.search(rf"^{re.escape(name)}\s*=\s*(\S+)", text, re.M)
    return float(m.group(1)) if m else None


def run(cmd):
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=str(ROOT))
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError(f"{cmd[0]} failed: {r.stdout[-500:]} {r.stderr[-1500:]}")
    rows = [ln.split() for ln in r.stdout.splitlines() if re.match(r"^\s*\d\.\d+e[-+]\d+\s", ln)]
    return {"wall_s": round(dt, 4), "steps": stat(r.stdout, "Steps"), "ls_iters": stat(r.stdout, "LS iters"),
            "rhs_evals": stat(r.stdout, "Implicit RHS fn evals"), "last_row": rows[-1] if rows else None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--tf-short", default="1e-6")
    ap.add_argument("--tf-long", default="1e-4")
    a = ap.parse_args()
    common = ["--nx", str(a.n), "--ny", str(a.n), "--nout", "1"]
    ref = str(ROOT / "oracle" / "_ref" / "bin" / "arkode_diffusion_2D_refcuda")
    ours = [sys.executable, str(ROOT / "apps" / "diffusion_2D" / "run.py")]
    out = {"mesh": f"{a.n}^2", "arms": {}}
    for name, cmd in (("reference_cuda", [ref]), ("b200", ours)):
        run(cmd + common + ["--tf", a.tf_short])  # warm-up: page the binaries in
        s = run(cmd + common + ["--tf", a.tf_short])
        l = run(cmd + common + ["--tf", a.tf_long])
        d_it = l["ls_iters"] - s["ls_iters"]
        out["arms"][name] = {"short": s, "long": l, "solve_s": round(l["wall_s"] - s["wall_s"], 4),
                             "ms_per_ls_iter": round((l["wall_s"] - s["wall_s"]) / d_it * 1e3, 4) if d_it else None}
    r, b = out["arms"]["reference_cuda"], out["arms"]["b200"]
    out["speedup_solve"] = round(r["solve_s"] / b["solve_s"], 3)
    if r["ms_per_ls_iter"] and b["ms_per_ls_iter"]:
        out["speedup_per_ls_iter"] = round(r["ms_per_ls_iter"] / b["ms_per_ls_iter"], 3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
