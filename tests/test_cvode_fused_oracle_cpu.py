"""Pins oracle/cvfused_oracle.py (the formulas the CUDA functors of b200vec_cvfused.cu are written from)
against the reference's own CPU implementation of CVODE's fused-kernel plugin, src/cvode/cvode_fused_stubs.c
on nvector_serial (baseline/_ref/lib/libsundials_cvode_fused_stubs.so), bit for bit, over the scalar
branches of N_VLinearSum the op sequences can reach.  Also: the plugin library exports every symbol
include/cvode_fused_b200.h declares.  No GPU."""
import ctypes as C
import re
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
import cvfused_oracle as orc  # noqa: E402

STUBS = ROOT / "baseline" / "_ref" / "lib" / "libsundials_cvode_fused_stubs.so"
V, D, I = C.c_void_p, C.c_double, C.c_int
N = 2053


@pytest.fixture(scope="module")
def stubs(refserial):
    if not STUBS.exists():
        pytest.skip("baseline/_ref not built (needs /root/reference)")
    from test_cvode_fused_gpu import _bind

    return refserial, _bind(C.CDLL(str(STUBS)))


def _vec(ref, a):
    return ref.L.N_VMake_Serial(len(a), a.ctypes.data_as(C.POINTER(C.c_double)), ref.ctx)


def _same(a, b):
    return np.array_equal(np.asarray(a).view(np.uint64), np.asarray(b).view(np.uint64))


def test_plugin_library_exports_the_declared_symbols():
    txt = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "cvode_fused_b200.h").read_text(), flags=re.S)
    names = sorted(set(re.findall(r"\b(cv\w+)\s*\(", txt)))
    assert len(names) == 7, names
    from sundials_b200 import _lib

    _lib.load()  # the main library first (the plugin's dependency, found through its rpath anyway)
    lib = C.CDLL(str(ROOT / "sundials_b200" / "lib" / "libsundials_cvode_fused_b200.so"))
    for n in names:
        assert hasattr(lib, n), n


@pytest.mark.parametrize("vec", [False, True])
def test_ewt(stubs, vec):
    ref, S = stubs
    r = np.random.default_rng(1)
    y, atol = r.uniform(-3, 3, N), r.uniform(1e-9, 1e-3, N)
    t, w = np.zeros(N), np.zeros(N)
    if vec:
        assert S.cvEwtSetSV_fused(1, 1e-4, _vec(ref, atol), _vec(ref, y), _vec(ref, t), _vec(ref, w)) == 0
        et, ew = orc.ewt(1e-4, atol, y)
    else:
        assert S.cvEwtSetSS_fused(1, 1e-4, 1e-6, _vec(ref, y), _vec(ref, t), _vec(ref, w)) == 0
        et, ew = orc.ewt(1e-4, 1e-6, y)
    assert _same(t, et) and _same(w, ew)


def test_constraints(stubs):
    ref, S = stubs
    r = np.random.default_rng(2)
    c, e, y, mm = r.choice([-2.0, -1.0, 0.0, 1.0, 2.0], N), r.uniform(10, 1e4, N), r.uniform(-1, 1, N), r.choice([0.0, 1.0], N)
    out = np.zeros(N)
    assert S.cvCheckConstraints_fused(_vec(ref, c), _vec(ref, e), _vec(ref, y), _vec(ref, mm), _vec(ref, out)) == 0
    assert _same(out, orc.constraints(c, e, y, mm))


@pytest.mark.parametrize("rl1,ngamma", [(0.37, -0.013), (1.0, -1.0), (-1.0, 1.0), (1.0, 0.25), (-1.0, -1.0)])
def test_nls_resid(stubs, rl1, ngamma):
    ref, S = stubs
    r = np.random.default_rng(3)
    z, yc, f = r.uniform(-1, 1, N), r.uniform(-1e-3, 1e-3, N), r.uniform(-50, 50, N)
    out = np.zeros(N)
    assert S.cvNlsResid_fused(rl1, ngamma, _vec(ref, z), _vec(ref, yc), _vec(ref, f), _vec(ref, out)) == 0
    assert _same(out, orc.nls_resid(rl1, ngamma, z, yc, f))


@pytest.mark.parametrize("h,rr", [(0.02, 0.05), (1.0, 1.0), (-1.0, -1.0), (1.0, 0.1), (-1.0, 0.3)])
def test_diag_form_y(stubs, h, rr):
    ref, S = stubs
    r = np.random.default_rng(4)
    fp, z, yp = r.uniform(-9, 9, N), r.uniform(-1, 1, N), r.uniform(-2, 2, N)
    ft, y = np.zeros(N), np.zeros(N)
    assert S.cvDiagSetup_formY(h, rr, _vec(ref, fp), _vec(ref, z), _vec(ref, yp), _vec(ref, ft), _vec(ref, y)) == 0
    eft, ey = orc.diag_form_y(h, rr, fp, z, yp)
    assert _same(ft, eft) and _same(y, ey)


@pytest.mark.parametrize("h", [0.0137, 0.1, -0.1, 1.0, -1.0])
def test_diag_build_m(stubs, h):
    ref, S = stubs
    r = np.random.default_rng(5)
    ft = r.uniform(-1e-2, 1e-2, N)
    ft[::40] = 0.0
    ft[1::57] = 1e-30
    fp, e, M0 = r.uniform(-5, 5, N), r.uniform(1, 1e6, N), r.uniform(-5, 5, N)
    bit, bc, y, M = np.zeros(N), np.zeros(N), np.zeros(N), M0.copy()
    u = float(np.finfo(np.float64).eps)
    assert S.cvDiagSetup_buildM(0.1, u, h, _vec(ref, ft), _vec(ref, fp), _vec(ref, e), _vec(ref, bit), _vec(ref, bc),
                                _vec(ref, y), _vec(ref, M)) == 0
    ebit, ebc, ey, eM = orc.diag_build_m(u, h, ft, fp, e, M0)
    assert 0 < np.count_nonzero(bit == 0.0) < N
    assert _same(bit, ebit) and _same(bc, ebc) and _same(y, ey) and _same(M, eM)


@pytest.mark.parametrize("rr", [0.83, 1.0, -1.0])
def test_diag_update_m(stubs, rr):
    ref, S = stubs
    r = np.random.default_rng(6)
    M0 = r.uniform(0.1, 3, N) * r.choice([-1.0, 1.0], N)
    M = M0.copy()
    assert S.cvDiagSolve_updateM(rr, _vec(ref, M)) == 0
    assert _same(M, orc.diag_update_m(rr, M0))
