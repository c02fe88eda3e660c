"""Build libsundials_nvecb200.so in-tree (sundials_b200/lib/).

nvcc cross-compiles the CUDA kernels for sm_100a (no GPU needed); gcc compiles
the C host code that fills the N_Vector ops table.  The result is one shared
library exporting the C ABI of include/b200vec.h and include/nvector_b200.h.

  python -m sundials_b200.build [--force] [--verbose]

The C host layer needs the SUNDIALS public headers (struct layouts of
N_Vector / N_Vector_Ops): by default the reference tree's include/ plus the
configuration header generated into baseline/_ref/include by baseline/Makefile;
override with SUNDIALS_INCLUDE="dir1:dir2" to build against an installed
SUNDIALS.  If no headers are available (e.g. on the GPU box) and a prebuilt
library exists, the prebuilt library is used as is.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "sundials_b200" / "csrc"
INC = ROOT / "include"
OBJ = ROOT / "build" / "obj"
LIBDIR = ROOT / "sundials_b200" / "lib"
LIB = LIBDIR / "libsundials_nvecb200.so"

CU_SOURCES = ["b200vec_ctx.cu", "b200vec_stream.cu", "b200vec_reduce.cu", "b200vec_fused.cu", "b200vec_cvfused.cu",
              "b200vec_comm.cu"]
C_SOURCES = ["nvector_b200.c", "sundials_iterative_b200.c"]
# separate tiny library: symbol interposition of SUNClassicalGS (see gs_interpose.c)
GS_LIB = LIBDIR / "libsundials_b200gs.so"
GS_SOURCES = ["gs_interpose.c"]
# separate library: CVODE's fused-kernel plugin boundary (takes the place of libsundials_cvode_fused_cuda)
CVF_LIB = LIBDIR / "libsundials_cvode_fused_b200.so"
CVF_SOURCES = ["cvode_fused_b200.c"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",              # bit-parity with nvector_serial: no FMA contraction
    "-Xcompiler", "-fPIC,-O2,-fvisibility=default",
    "-Xptxas", "-v",
]
GCC = "/usr/bin/gcc" if Path("/usr/bin/gcc").exists() else "gcc"


def sundials_includes() -> list[str]:
    env = os.environ.get("SUNDIALS_INCLUDE")
    if env:
        return [d for d in env.split(":") if d]
    cands = [Path("/root/reference/include"), ROOT / "baseline" / "_ref" / "include"]
    return [str(c) for c in cands if c.exists()]


def have_sundials_headers() -> bool:
    incs = sundials_includes()
    return any((Path(d) / "sundials" / "sundials_nvector.h").exists() for d in incs) and any(
        (Path(d) / "sundials" / "sundials_config.h").exists() for d in incs
    )


def _digest(paths: list[Path], extra: str) -> str:
    h = hashlib.sha256(extra.encode())
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def _run(cmd: list[str], verbose: bool) -> str:
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"build step failed: {' '.join(cmd[:4])} ...")
    return r.stdout + r.stderr


def build(force: bool = False, verbose: bool = False) -> Path:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    headers = list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + list(INC.glob("*.h"))
    sources = [CSRC / s for s in CU_SOURCES + C_SOURCES + GS_SOURCES + CVF_SOURCES]
    stamp = LIBDIR / ".build_digest"
    digest = _digest(headers + sources, " ".join(NVCC_FLAGS))
    if not force and LIB.exists() and GS_LIB.exists() and CVF_LIB.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB
    if not Path(nvcc).exists() or not have_sundials_headers():
        if LIB.exists():
            # GPU box: no reference headers / nothing to rebuild -- use the prebuilt library
            return LIB
        raise RuntimeError(
            "cannot build libsundials_nvecb200.so: need nvcc and the SUNDIALS headers "
            "(run `make -C baseline` first, or set SUNDIALS_INCLUDE)"
        )
    OBJ.mkdir(parents=True, exist_ok=True)
    LIBDIR.mkdir(parents=True, exist_ok=True)
    jobs = []
    for s in CU_SOURCES:
        o = OBJ / (s + ".o")
        jobs.append((o, [nvcc, *NVCC_FLAGS, f"-I{INC}", f"-I{CSRC}", "-c", str(CSRC / s), "-o", str(o)]))
    sun_inc = [f"-I{d}" for d in sundials_includes()]
    for s in C_SOURCES:
        o = OBJ / (s + ".o")
        jobs.append((o, [GCC, "-O2", "-std=gnu99", "-fPIC", "-Wall", f"-I{INC}", *sun_inc, "-c", str(CSRC / s),
                         "-o", str(o)]))
    logs = []
    with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
        for out in ex.map(lambda j: _run(j[1], verbose), jobs):
            logs.append(out)
    (ROOT / "build" / "ptxas.log").write_text("\n".join(logs))
    objs = [str(o) for o, _ in jobs]
    cuda_lib = Path(nvcc).resolve().parent.parent / "lib64"
    gxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    # host link only (no relocatable device code): keeps the library pure sm_100a
    _run([gxx, "-shared", "-o", str(LIB), *objs, f"-L{cuda_lib}", "-lcudart_static", "-Wl,--no-undefined", "-ldl",
          "-lm", "-lpthread", "-lrt"], verbose)
    # the interposer resolves SUNClassicalGS_B200 / N_VCloneEmpty_B200 from the main library
    _run([GCC, "-O2", "-std=gnu99", "-fPIC", "-shared", "-Wall", f"-I{INC}", *sun_inc, "-o", str(GS_LIB),
          *[str(CSRC / s) for s in GS_SOURCES], f"-L{LIBDIR}", "-lsundials_nvecb200", "-Wl,-rpath,$ORIGIN", "-ldl"],
         verbose)
    # the fused-kernel plugin: seven reference symbols over the b200vec_cv_* kernels of the main library
    _run([GCC, "-O2", "-std=gnu99", "-fPIC", "-shared", "-Wall", f"-I{INC}", *sun_inc, "-o", str(CVF_LIB),
          *[str(CSRC / s) for s in CVF_SOURCES], f"-L{LIBDIR}", "-lsundials_nvecb200", "-Wl,-rpath,$ORIGIN"],
         verbose)
    stamp.write_text(digest)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
