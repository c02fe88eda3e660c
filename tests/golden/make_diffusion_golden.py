"""Generate tests/golden/diffusion_2D/*.out by RUNNING THE REFERENCE HERE: the
unmodified benchmarks/diffusion_2D (main_arkode.cpp + mpi_serial backend +
nvector_parallel), compiled by path against the single-rank MPI stand-in
(tests/c/shim_mpi/mpi.h) into oracle/_ref/bin/arkode_diffusion_2D_ref.

    make -C oracle ref && make -C tests/c && python tests/golden/make_diffusion_golden.py
"""
import json
import subprocess
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
EXE = ROOT / "oracle" / "_ref" / "bin" / "arkode_diffusion_2D_ref"
OUT = Path(__file__).resolve().parent / "diffusion_2D"

CASES = {
    "default_32x32": [],
    "64x64": ["--nx", "64", "--ny", "64"],
    "128x96_tf0.2": ["--nx", "128", "--ny", "96", "--tf", "0.2", "--nout", "4"],
    "33x31_scalar_path": ["--nx", "33", "--ny", "31"],
    "64x64_noforcing": ["--nx", "64", "--ny", "64", "--noforcing"],
    "64x64_gmres": ["--nx", "64", "--ny", "64", "--ls", "gmres"],
    "256x256_tf0.1": ["--nx", "256", "--ny", "256", "--tf", "0.1", "--nout", "2"],
}


def main():
    OUT.mkdir(exist_ok=True)
    manifest = {}
    for tag, args in CASES.items():
        t0 = time.time()
        r = subprocess.run([str(EXE), *args], capture_output=True, text=True)
        dt = time.time() - t0
        (OUT / f"{tag}.out").write_text(r.stdout)
        manifest[tag] = {"args": args, "returncode": r.returncode, "reference_cpu_seconds": round(dt, 3)}
        print(tag, manifest[tag])
    (OUT / "MANIFEST.json").write_text(json.dumps(manifest, indent=1, sort_keys=True) + "\n")


if __name__ == "__main__":
    main()
