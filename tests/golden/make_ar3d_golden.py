"""Generate tests/golden/advection_reaction_3D/ by RUNNING THE REFERENCE HERE.

  rhs_*.npz   inputs + outputs of the reference's own SetIC / Advection / Reaction /
              AdvectionReaction / SolveReactionLinSys (benchmarks/advection_reaction_3D/raja/
              rhs3D.hpp, advection_reaction_3D.cpp) driven by tests/c/ar3d_rhs_dump.cpp
  *.out       stdout of the unmodified benchmark (oracle/_ref/bin/advection_reaction_3D_ref:
              RAJA sequential + single-rank MPI stand-ins, nvector_serial under MPIPlusX)
  *.final.npz the last line of its u/v/w.000000.txt solution files (%.16e: exact doubles)

    make -C oracle ref && make -C tests/c && python tests/golden/make_ar3d_golden.py
"""
import json
import subprocess
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
BIN = ROOT / "oracle" / "_ref" / "bin"
OUT = Path(__file__).resolve().parent / "advection_reaction_3D"

RHS_CASES = {
    # tag: (npts, c, gamma, seed)
    "rhs_8_cpos": (8, 0.01, 1.3e-3, 11),
    "rhs_8_cneg": (8, -0.01, 2.0e-2, 12),
    "rhs_6_cpos": (6, 0.5, 1.0e-4, 13),
    "rhs_5_cneg": (5, -0.25, 5.0e-3, 14),
}

RUN_CASES = {
    "dirk_newton_8": ["--npts", "8", "--tf", "1.0", "--nout", "5"],
    "dirk_newton_nopre_8": ["--npts", "8", "--tf", "0.5", "--nout", "2", "--nopre"],
    "dirk_fixedpoint_8": ["--npts", "8", "--tf", "0.01", "--nout", "2", "--nls", "fixedpoint"],
    "imex_newton_8": ["--npts", "8", "--tf", "1.0", "--nout", "5", "--method", "ARK-IMEX"],
    # (--nls tl-newton cannot be run from the reference here: its TaskLocalNewton constructor
    #  dereferences the integer SUNComm of a non-MPI SUNDIALS build as a pointer,
    #  arkode_driver.cpp:762-770 -- the re-host's tl-newton is checked against imex_newton_8)
    "imex_fixedpoint_8": ["--npts", "8", "--tf", "0.01", "--nout", "2", "--method", "ARK-IMEX", "--nls", "fixedpoint"],
    "imex_newton_fused_8": ["--npts", "8", "--tf", "1.0", "--nout", "5", "--method", "ARK-IMEX", "--fused"],
    "erk_8": ["--npts", "8", "--tf", "0.001", "--nout", "2", "--method", "ERK"],
    "bdf_newton_8": ["--npts", "8", "--tf", "1.0", "--nout", "5", "--method", "CV-BDF"],
    "bdf_fixedpoint_8": ["--npts", "8", "--tf", "0.01", "--nout", "2", "--method", "CV-BDF", "--nls", "fixedpoint"],
    "adams_8": ["--npts", "8", "--tf", "0.01", "--nout", "2", "--method", "CV-ADAMS"],
    "imex_newton_10_generic": ["--npts", "10", "--tf", "1.0", "--nout", "4", "--method", "ARK-IMEX"],
    "dirk_newton_cneg_8": ["--npts", "8", "--tf", "0.5", "--nout", "2", "--c", "-0.05"],
    "imex_newton_16": ["--npts", "16", "--tf", "1.0", "--nout", "5", "--method", "ARK-IMEX"],
    "dirk_newton_24_order4": ["--npts", "24", "--tf", "0.5", "--nout", "2", "--order", "4"],
}


def rhs_case(tag, npts, c, gamma, seed):
    rng = np.random.default_rng(seed)
    n = npts ** 3
    y = np.empty((n, 3))
    # near the steady state (u, v, w) = (1, 3.5, 3) so that I - gamma J is well conditioned
    y[:, 0] = 1.0 + 0.2 * rng.uniform(-1, 1, n)
    y[:, 1] = 3.5 + 0.2 * rng.uniform(-1, 1, n)
    y[:, 2] = 3.0 + 0.2 * rng.uniform(-1, 1, n)
    b = rng.uniform(-1, 1, (n, 3))
    with tempfile.TemporaryDirectory() as td:
        fin, fout = Path(td) / "in.bin", Path(td) / "out.bin"
        np.concatenate([y.ravel(), b.ravel()]).tofile(fin)
        r = subprocess.run([str(BIN / "ar3d_rhs_dump"), str(npts), repr(c), repr(gamma), str(fin), str(fout)],
                           capture_output=True, text=True, cwd=td)
        assert r.returncode == 0, r.stdout + r.stderr
        o = np.fromfile(fout).reshape(5, n * 3)
    np.savez_compressed(OUT / f"{tag}.npz", npts=npts, c=c, gamma=gamma, y=y.ravel(), b=b.ravel(), ic=o[0], fe=o[1],
                        fi=o[2], f=o[3], x=o[4])
    print(tag, "ok")


def run_case(tag, args):
    with tempfile.TemporaryDirectory() as td:
        t0 = time.time()
        r = subprocess.run([str(BIN / "advection_reaction_3D_ref"), *args, "--output-dir", td], capture_output=True,
                           text=True, cwd=td)
        dt = time.time() - t0
        (OUT / f"{tag}.out").write_text(r.stdout.replace(td, "."))
        last = {}
        for s in "uvw":
            lines = (Path(td) / f"{s}.000000.txt").read_text().splitlines()
            last[s] = np.array(lines[-1].split(), dtype=np.float64)
        np.savez_compressed(OUT / f"{tag}.final.npz", **last)
    print(tag, r.returncode, round(dt, 2))
    return {"args": args, "returncode": r.returncode, "reference_cpu_seconds": round(dt, 3)}


def main():
    OUT.mkdir(exist_ok=True)
    for tag, (npts, c, gamma, seed) in RHS_CASES.items():
        rhs_case(tag, npts, c, gamma, seed)
    manifest = {tag: run_case(tag, args) for tag, args in RUN_CASES.items()}
    (OUT / "MANIFEST.json").write_text(json.dumps(manifest, indent=1, sort_keys=True) + "\n")


if __name__ == "__main__":
    main()
