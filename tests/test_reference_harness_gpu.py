"""The reference's OWN tests and examples, compiled by path and unmodified,
running on NVECTOR_B200 through the N_Vector_Ops table (tests/c/Makefile):

  * test_nvector.c  -- the reference's known-answer harness for every op,
    all three memory kinds, fused ops disabled (generic fallback) and enabled;
  * SUNModifiedGS / SUNClassicalGS (sundials_iterative.c) serial vs B200;
  * SPGMR / SPFGMR / PCG unit tests (1e-13 solves, both Gram-Schmidt types);
  * CVODE cvDiurnal_kry, ARKODE ark_heat1D + ark_heat2D, IDA idaHeat2D_kry,
    KINSOL kinFoodWeb_kry + kinLaplace_picard_kry.

The integrator programs must print output BYTE-IDENTICAL to the goldens that the
same sources produced on the reference's nvector_serial
(tests/golden/examples/*.out, tests/golden/make_example_golden.py): same
solution digits, same step / iteration / failure counters.  That holds because
streaming ops are bit-exact and, at these problem sizes (N <= 1024), reductions
take the exact-order path.
"""
import json
import subprocess
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "oracle" / "_ref" / "bin"      # programs that contain the reference's CPU vectors
BINB = ROOT / "baseline" / "_ref" / "bin"   # reference programs on NVECTOR_B200 + the host framework only
GOLD = ROOT / "tests" / "golden" / "examples"
MANIFEST = json.loads((GOLD / "MANIFEST.json").read_text())


def _run(exe, *args, timeout=600):
    p = (BINB if (BINB / exe).exists() else BIN) / exe
    assert p.exists(), f"{p} missing: run `make -C tests/c` where /root/reference exists"
    return subprocess.run([str(p), *map(str, args)], capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("length", [1000, 100_000])
def test_reference_nvector_unit_harness(length):
    r = _run("test_nvector_b200", length, 0)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "SUCCESS: NVECTOR_B200 module passed all tests" in r.stdout
    assert "FAILED" not in r.stdout
    # every op of the table was exercised (3 memory kinds x PASSED lines)
    assert r.stdout.count("PASSED test -- N_VLinearCombination Case") >= 3 * 2 * 3


def test_reference_nvector_unit_harness_tiny_lengths():
    # the reference CUDA driver's CTest lengths (cuda/CMakeLists.txt:24-26) start at 3
    for length in (7, 33):
        r = _run("test_nvector_b200", length, 0, "device")
        assert r.returncode == 0 and "SUCCESS" in r.stdout, r.stdout[-2000:]


def test_gram_schmidt_bit_identical_small():
    r = _run("test_gs_b200", 1000, 10, 0)
    assert r.returncode == 0, r.stdout
    assert "SUCCESS" in r.stdout and "MISMATCH" not in r.stdout


def test_gram_schmidt_tolerance_large():
    r = _run("test_gs_b200", 300_000, 20, 1e-13)
    assert r.returncode == 0, r.stdout
    assert "SUCCESS" in r.stdout


def test_fused_gs_matches_reference_gs_on_serial():
    """third / fourth pass of test_gs_b200: SUNClassicalGS_B200 (2 kernels per column) and SUNModifiedGS_B200
    (k + 1 kernels per column) on NVECTOR_B200 against the reference's SUNClassicalGS / SUNModifiedGS on
    nvector_serial -- bit-identical at n = 1000, 1e-13 at 300 000"""
    r = _run("test_gs_b200", 1000, 10, 0)
    assert r.returncode == 0 and "MISMATCH" not in r.stdout, r.stdout
    assert r.stdout.count("fused-cgs k=") == 10 and r.stdout.count("fused-mgs k=") == 10, r.stdout
    r = _run("test_gs_b200", 300_000, 20, 1e-13)
    assert r.returncode == 0 and "MISMATCH" not in r.stdout, r.stdout
    assert r.stdout.count("fused-cgs k=") == 20 and r.stdout.count("fused-mgs k=") == 20, r.stdout


@pytest.mark.parametrize("prog", ["test_sunlinsol_spgmr_b200", "test_sunlinsol_spfgmr_b200"])
@pytest.mark.parametrize("gstype,routine", [(1, "SUNModifiedGS_B200"), (2, "SUNClassicalGS_B200")])
def test_krylov_solvers_use_the_fused_gs_by_symbol_interposition(prog, gstype, routine):
    """the reference's UNMODIFIED SPGMR / SPFGMR unit tests (both Gram-Schmidt types) with
    libsundials_b200gs.so preloaded: their SUNModifiedGS / SUNClassicalGS calls land in the fused _B200
    routines (interposed, reference unmodified) and the 1e-13 solves still pass, printing the same text"""
    import os

    so = ROOT / "sundials_b200" / "lib" / "libsundials_b200gs.so"
    assert so.exists()
    env = dict(os.environ, LD_PRELOAD=str(so), B200GS_REPORT="1")
    p = BINB / prog
    # args of the reference CTests (spgmr/serial/CMakeLists.txt:34-37: n, gstype, pretype, maxl, tol, timing;
    # spfgmr/serial/CMakeLists.txt:34-35: n, gstype, maxl, tol, timing)
    args = [str(p), "100", str(gstype)] + (["1"] if "spgmr" in prog else []) + ["100", "1e-13", "0"]
    r = subprocess.run(args, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "FAIL" not in r.stdout
    calls = [int(x.rsplit(":", 1)[1]) for x in r.stderr.splitlines() if f"{routine} calls" in x]
    assert calls and calls[0] > 0, r.stderr[-500:]
    # and without the preload the same program never reaches it, and prints the same solve
    r0 = subprocess.run(args, capture_output=True, text=True, timeout=600)
    assert r0.returncode == 0 and r0.stdout == r.stdout


def test_reference_cuda_example_with_device_rhs_kernels():
    """examples/cvode/cuda/cvAdvDiff_kry_cuda.cu, UNMODIFIED (shim header maps N_V*_Cuda -> N_V*_B200):
    RHS / Jv are user CUDA kernels on the legacy default stream reading N_VGetDeviceArrayPointer, the
    vector ops run on a user stream -- the only reference program that exercises NVECTOR_B200's
    device mode with foreign kernels.  Every printed norm and every counter must equal the reference's
    own committed output of the program on nvector_cuda (nst = 143, nfe = 206, nni = 203, nli = 225,
    netf = 2).  Only the integer-workspace sizes leniw / leniwLS differ: N_VSpace reports liw = 1 per
    vector like nvector_serial (serial:362-373), nvector_cuda reports 2 (cuda:749)."""
    import re

    r = _run("cvAdvDiff_kry_cuda_b200")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    want = (GOLD / "cvAdvDiff_kry_cuda.refcuda.out").read_text()
    strip = lambda t: re.sub(r"(leniw(LS)?\s*=\s*)\d+", r"\1#", t)  # noqa: E731
    assert strip(r.stdout) == strip(want), _first_diff(strip(r.stdout), strip(want))
    assert "nst     =   143" in r.stdout and "nli     =   225" in r.stdout


@pytest.mark.parametrize("tag", sorted(MANIFEST))
def test_reference_program_output_identical_to_serial_golden(tag):
    e = MANIFEST[tag]
    r = _run(e["program"] + "_b200", *e["args"])
    want = (GOLD / f"{tag}.out").read_text()
    assert r.returncode == e["returncode"], r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout == want, _first_diff(r.stdout, want)


def _first_diff(a, b):
    for i, (x, y) in enumerate(zip(a.splitlines(), b.splitlines())):
        if x != y:
            return f"line {i + 1}:\n  b200  : {x}\n  serial: {y}"
    return f"length differs: {len(a)} vs {len(b)}"
