"""CPU checks of the bench's C suite driver (apps/nvector_perf, the re-host of
benchmarks/nvector/test_nvector_performance.c) on the reference's own CPU vectors: the op list and its
byte model, every op reachable through the ops table, identical results on nvector_serial and
nvector_openmp, and the reference arm of bench.py end to end (rank 0 prints one JSON line)."""
import ctypes as C
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


@pytest.fixture(scope="module")
def bench():
    import bench as b

    if not (ROOT / "oracle" / "_ref" / "lib" / "libsundials_ref.so").exists():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return b


def _suite(b, n, threads):
    lib, perf = b.load_reference(), b.load_perf()
    sctx = C.c_void_p()
    assert lib.SUNContext_Create(0, C.byref(sctx)) == 0
    vec = b.cpu_vectors(lib, sctx, n, threads)
    return lib, vec, b.Suite(perf, vec)


def test_op_list_and_byte_model(bench):
    lib, vec, s = _suite(bench, 1000, 1)
    assert s.nops == 55 and len(set(s.names)) == 55
    # SURVEY section 8d byte model, nvecs = 8, nsums = 4: 2952 algorithmic bytes per element and step
    assert s.bytes_per_elt_step == 2952
    by = dict(zip(s.names, s.bpe))
    assert by["N_VLinearSum-9"] == 24 and by["N_VConst"] == 8 and by["N_VDotProd"] == 16
    assert by["N_VLinearCombination-3"] == 72 and by["N_VScaleAddMulti-2"] == 136 and by["N_VDotProdMulti"] == 72
    assert by["N_VScaleAddMultiVectorArray"] == 576 and by["N_VLinearCombinationVectorArray"] == 320
    assert sum(s.scalar) == 18          # ops that hand scalars back to the host
    bench.free_vectors(lib, vec)


def test_same_results_on_serial_and_openmp(bench):
    n = 20_000
    lib, v1, s1 = _suite(bench, n, 1)
    _, v4, s4 = _suite(bench, n, 4)
    for s in (s1, s4):
        s.step(2)
        s.check()
    for i, name in enumerate(s1.names):
        if not s1.scalar[i]:
            continue
        a, b = s1.perf.nvperf_result(s1.h, i), s4.perf.nvperf_result(s4.h, i)
        assert abs(a - b) <= 1e-13 * max(abs(a), 1.0) * 50, (name, a, b)
    # a streaming output, bit for bit
    import numpy as np

    z1 = np.ctypeslib.as_array(lib.N_VGetArrayPointer(v1["Z"][1]), shape=(n,))
    z4 = np.ctypeslib.as_array(lib.N_VGetArrayPointer(v4["Z"][1]), shape=(n,))
    assert np.array_equal(z1.view(np.uint64), z4.view(np.uint64))
    bench.free_vectors(lib, v1)
    bench.free_vectors(lib, v4)


def test_reference_arm_prints_one_contract_line(bench):
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--log2n", "16", "--steps", "3",
                        "--warmup", "3", "--gpus", "1"], capture_output=True, text=True, timeout=300, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GB/s" and d["higher_is_better"] is True
    assert d["gpus_used"] == 0 and d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["config"]["length_per_gpu"] == 1 << 16 and "nvector_openmp" in d["cpu_baseline"]["sample"] or \
        "nvector_serial" in d["cpu_baseline"]["sample"]
    assert d["result_checksum"] > 0
