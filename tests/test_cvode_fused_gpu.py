"""libsundials_cvode_fused_b200.so against the reference's own CPU implementation of the same plugin
boundary (SURVEY §8 row f-N3, include/cvode_fused_b200.h).

Oracle: baseline/_ref/lib/libsundials_cvode_fused_stubs.so = src/cvode/cvode_fused_stubs.c, unmodified,
running on nvector_serial.  Each of the seven functions is called on both sides with the same seeded
inputs; every vector the function may write is compared BIT FOR BIT, and so is the return value
(including the atolmin0 failure of the error-weight functions, after which the weights must be
untouched).  The scalar arguments sweep the branches of N_VLinearSum's case analysis the sequences can
reach (+-1, a == b, a == -b, general).
"""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
STUBS = ROOT / "baseline" / "_ref" / "lib" / "libsundials_cvode_fused_stubs.so"
FUSED = ROOT / "sundials_b200" / "lib" / "libsundials_cvode_fused_b200.so"

V, D, I = C.c_void_p, C.c_double, C.c_int
SIGS = {
    "cvEwtSetSS_fused": [I, D, D, V, V, V],
    "cvEwtSetSV_fused": [I, D, V, V, V, V],
    "cvCheckConstraints_fused": [V, V, V, V, V],
    "cvNlsResid_fused": [D, D, V, V, V, V],
    "cvDiagSetup_formY": [D, D, V, V, V, V, V],
    "cvDiagSetup_buildM": [D, D, D, V, V, V, V, V, V, V],
    "cvDiagSolve_updateM": [D, V],
}
LENGTHS = [1, 7, 1000, 4099, 300_001, (1 << 20) + 5]


def _bind(lib):
    for name, args in SIGS.items():
        f = getattr(lib, name)
        f.restype, f.argtypes = I, args
    return lib


@pytest.fixture(scope="module")
def sides():
    from _oracle import RefSerial
    from sundials_b200.plugin import B200Plugin

    ref = RefSerial()  # loads libsundials_ref.so + the host framework (RTLD_GLOBAL)
    assert STUBS.exists(), f"{STUBS} missing: make -C baseline"
    assert FUSED.exists(), f"{FUSED} missing: python -m sundials_b200.build"
    P = B200Plugin()
    return ref, _bind(C.CDLL(str(STUBS))), P, _bind(C.CDLL(str(FUSED)))


class _Pair:
    """the same named arrays as nvector_serial vectors (wrapping the numpy arrays) and as NVECTOR_B200 vectors"""

    def __init__(self, sides, arrays, kind=0):
        self.ref, self.stubs, self.P, self.fused = sides
        self.cpu = {k: np.array(a, dtype=np.float64) for k, a in arrays.items()}
        self.keep = {k: a.ctypes.data_as(C.POINTER(C.c_double)) for k, a in self.cpu.items()}
        self.s = {k: self.ref.L.N_VMake_Serial(len(a), self.keep[k], self.ref.ctx) for k, a in self.cpu.items()}
        self.b = {}
        for k, a in arrays.items():
            h = self.P.new(len(a), None, kind, fused=True)
            self.P.host(h, len(a))[...] = a
            self.P.to_device(h)
            self.b[k] = h

    def gpu(self, k):
        h = self.b[k]
        self.P.from_device(h)
        return np.array(self.P.host(h, len(self.cpu[k])))

    def check(self, *names):
        for k in names:
            got, want = self.gpu(k), self.cpu[k]
            bad = np.flatnonzero(got.view(np.uint64) != want.view(np.uint64))
            assert bad.size == 0, f"{k}: {bad.size} elements differ, first at {bad[0]}: {got[bad[0]]!r} vs {want[bad[0]]!r}"

    def close(self):
        for h in self.b.values():
            self.P.Destroy(h)
        for v in self.s.values():
            self.ref.L.N_VDestroy_Serial(v)


def _rng(n, seed):
    return np.random.default_rng(seed + n)


@pytest.mark.parametrize("n", LENGTHS)
@pytest.mark.parametrize("vector_atol", [False, True])
@pytest.mark.parametrize("atolmin0,zero", [(0, False), (1, False), (1, True)])
def test_ewt_set(sides, n, vector_atol, atolmin0, zero):
    r = _rng(n, 1)
    y = r.uniform(-3, 3, n)
    atol = r.uniform(1e-9, 1e-3, n)
    sab = 1e-6
    if zero:  # a zero denominator: y_i = 0 where atol_i = 0
        y[n // 2] = 0.0
        atol[n // 2] = 0.0
        sab = 0.0
    p = _Pair(sides, dict(y=y, atol=atol, tempv=r.uniform(1, 2, n), w=r.uniform(1, 2, n)))
    try:
        if vector_atol:
            rs = p.stubs.cvEwtSetSV_fused(atolmin0, 1e-4, p.s["atol"], p.s["y"], p.s["tempv"], p.s["w"])
            rb = p.fused.cvEwtSetSV_fused(atolmin0, 1e-4, p.b["atol"], p.b["y"], p.b["tempv"], p.b["w"])
        else:
            rs = p.stubs.cvEwtSetSS_fused(atolmin0, 1e-4, sab, p.s["y"], p.s["tempv"], p.s["w"])
            rb = p.fused.cvEwtSetSS_fused(atolmin0, 1e-4, sab, p.b["y"], p.b["tempv"], p.b["w"])
        assert rs == rb == (-1 if zero else 0)
        p.check("tempv", "w", "y", "atol")  # after a failure: tempv written, w untouched, on both sides
    finally:
        p.close()


@pytest.mark.parametrize("n", LENGTHS)
def test_check_constraints(sides, n):
    r = _rng(n, 2)
    p = _Pair(sides, dict(c=r.choice([-2.0, -1.0, 0.0, 1.0, 2.0], n), ewt=r.uniform(10, 1e4, n), y=r.uniform(-1, 1, n),
                          mm=r.choice([0.0, 1.0], n), tmp=np.zeros(n)))
    try:
        args = ("c", "ewt", "y", "mm", "tmp")
        assert p.stubs.cvCheckConstraints_fused(*[p.s[k] for k in args]) == 0
        assert p.fused.cvCheckConstraints_fused(*[p.b[k] for k in args]) == 0
        p.check("tmp", "c", "ewt", "y", "mm")
    finally:
        p.close()


@pytest.mark.parametrize("n", LENGTHS)
@pytest.mark.parametrize("rl1,ngamma", [(0.37, -0.013), (1.0, -1.0), (-1.0, 1.0), (1.0, 0.25)])
def test_nls_resid(sides, n, rl1, ngamma):
    r = _rng(n, 3)
    p = _Pair(sides, dict(zn1=r.uniform(-1, 1, n), ycor=r.uniform(-1e-3, 1e-3, n), ftemp=r.uniform(-50, 50, n),
                          res=np.zeros(n)))
    try:
        args = ("zn1", "ycor", "ftemp", "res")
        assert p.stubs.cvNlsResid_fused(rl1, ngamma, *[p.s[k] for k in args]) == 0
        assert p.fused.cvNlsResid_fused(rl1, ngamma, *[p.b[k] for k in args]) == 0
        p.check("res")
    finally:
        p.close()


@pytest.mark.parametrize("n", LENGTHS)
@pytest.mark.parametrize("h,rr", [(0.02, 0.05), (1.0, 1.0), (-1.0, -1.0), (1.0, 0.1)])
def test_diag_form_y(sides, n, h, rr):
    r = _rng(n, 4)
    p = _Pair(sides, dict(fpred=r.uniform(-9, 9, n), zn1=r.uniform(-1, 1, n), ypred=r.uniform(-2, 2, n), ftemp=np.zeros(n),
                          y=np.zeros(n)))
    try:
        args = ("fpred", "zn1", "ypred", "ftemp", "y")
        assert p.stubs.cvDiagSetup_formY(h, rr, *[p.s[k] for k in args]) == 0
        assert p.fused.cvDiagSetup_formY(h, rr, *[p.b[k] for k in args]) == 0
        p.check("ftemp", "y")
    finally:
        p.close()


@pytest.mark.parametrize("n", LENGTHS)
@pytest.mark.parametrize("h", [0.0137, 0.1, -0.1, 1.0, -1.0])  # general, a == -b, a == b, b == -1, b == +1
def test_diag_build_m(sides, n, h):
    r = _rng(n, 5)
    ftemp = r.uniform(-1e-2, 1e-2, n)
    ftemp[:: max(1, n // 50)] = 0.0  # round-off guard: bit = 0, bitcomp = -1 there
    ftemp[1 :: max(2, n // 37)] = 1e-30
    p = _Pair(sides, dict(ftemp=ftemp, fpred=r.uniform(-5, 5, n), ewt=r.uniform(1, 1e6, n), bit=np.zeros(n),
                          bitcomp=np.zeros(n), y=np.zeros(n), M=r.uniform(-5, 5, n)))
    try:
        uround = float(np.finfo(np.float64).eps)
        args = ("ftemp", "fpred", "ewt", "bit", "bitcomp", "y", "M")
        assert p.stubs.cvDiagSetup_buildM(0.1, uround, h, *[p.s[k] for k in args]) == 0
        assert p.fused.cvDiagSetup_buildM(0.1, uround, h, *[p.b[k] for k in args]) == 0
        assert np.count_nonzero(p.cpu["bit"] == 0.0) > 0  # the guard was exercised
        p.check("bit", "bitcomp", "y", "M", "ftemp", "fpred", "ewt")
    finally:
        p.close()


@pytest.mark.parametrize("n", LENGTHS)
@pytest.mark.parametrize("rr", [0.83, 1.0, -1.0])
def test_diag_update_m(sides, n, rr):
    r = _rng(n, 6)
    p = _Pair(sides, dict(M=r.uniform(0.1, 3, n) * r.choice([-1.0, 1.0], n)))
    try:
        assert p.stubs.cvDiagSolve_updateM(rr, p.s["M"]) == 0
        assert p.fused.cvDiagSolve_updateM(rr, p.b["M"]) == 0
        p.check("M")
    finally:
        p.close()


@pytest.mark.parametrize("kind", [1, 2])  # managed, pinned: results are host-visible when the call returns
def test_host_visible_memory_kinds_are_coherent_on_return(sides, kind):
    n = 50_001
    r = _rng(n, 7)
    p = _Pair(sides, dict(zn1=r.uniform(-1, 1, n), ycor=r.uniform(-1, 1, n), ftemp=r.uniform(-1, 1, n), res=np.zeros(n)), kind=kind)
    try:
        args = ("zn1", "ycor", "ftemp", "res")
        assert p.stubs.cvNlsResid_fused(0.5, -0.25, *[p.s[k] for k in args]) == 0
        assert p.fused.cvNlsResid_fused(0.5, -0.25, *[p.b[k] for k in args]) == 0
        got = np.array(p.P.host(p.b["res"], n))  # NO copy-back: the host pointer is the data
        assert np.array_equal(got.view(np.uint64), p.cpu["res"].view(np.uint64))
    finally:
        p.close()


def test_the_unfused_stubs_on_the_b200_vector_give_the_same_bits(sides):
    """the reference stubs driving NVECTOR_B200 through its ops table (what CVODE does with the fused
    kernels off) against the fused kernel: the claim 'fused on prints what fused off prints'"""
    n = 100_003
    r = _rng(n, 8)
    base = dict(ftemp=r.uniform(-1e-2, 1e-2, n), fpred=r.uniform(-5, 5, n), ewt=r.uniform(1, 1e6, n), bit=np.zeros(n),
                bitcomp=np.zeros(n), y=np.zeros(n), M=r.uniform(-5, 5, n))
    a, b = _Pair(sides, base), _Pair(sides, base)
    try:
        uround = float(np.finfo(np.float64).eps)
        args = ("ftemp", "fpred", "ewt", "bit", "bitcomp", "y", "M")
        assert a.stubs.cvDiagSetup_buildM(0.1, uround, 0.02, *[a.b[k] for k in args]) == 0   # 10 launches
        assert b.fused.cvDiagSetup_buildM(0.1, uround, 0.02, *[b.b[k] for k in args]) == 0   # 1 launch
        for k in ("bit", "bitcomp", "y", "M"):
            assert np.array_equal(a.gpu(k).view(np.uint64), b.gpu(k).view(np.uint64)), k
    finally:
        a.close()
        b.close()
