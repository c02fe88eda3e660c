"""numpy-facing bindings of the two CPU checkers (TEST INFRASTRUCTURE ONLY):

  * Oracle    -- oracle/_build/libnvec_oracle.so, our plain-C restatement of
                 nvector_serial (oracle/nvec_oracle.c);
  * RefSerial -- oracle/_ref/lib/libsundials_ref.so, the UNMODIFIED reference
                 nvector_serial compiled from /root/reference by oracle/Makefile
                 (present here and on the GPU box; built only where
                 /root/reference exists).

Both expose the same method names so a test can run one against the other.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_SO = ROOT / "oracle" / "_build" / "libnvec_oracle.so"
REF_SO = ROOT / "oracle" / "_ref" / "lib" / "libsundials_ref.so"

dp = C.POINTER(C.c_double)
dpp = C.POINTER(dp)


def build_oracle() -> Path:
    subprocess.run(["make", "-s", "-C", str(ROOT / "oracle"), "oracle"], check=True)
    return ORACLE_SO


def _p(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(dp)


def _tab(arrs):
    t = (dp * len(arrs))(*[_p(a) for a in arrs])
    return t


def _coef(c):
    return (C.c_double * len(c))(*[float(v) for v in c])


class Oracle:
    """oracle/nvec_oracle.c through ctypes; arrays are numpy float64, modified in place."""

    def __init__(self):
        if not ORACLE_SO.exists():
            build_oracle()
        L = C.CDLL(str(ORACLE_SO))
        self.L = L
        D, I, N = C.c_double, C.c_int, C.c_int64
        sig = {
            "orc_linear_sum": (None, [D, dp, D, dp, dp, N]),
            "orc_const": (None, [D, dp, N]),
            "orc_prod": (None, [dp, dp, dp, N]),
            "orc_div": (None, [dp, dp, dp, N]),
            "orc_scale": (None, [D, dp, dp, N]),
            "orc_abs": (None, [dp, dp, N]),
            "orc_inv": (None, [dp, dp, N]),
            "orc_add_const": (None, [dp, D, dp, N]),
            "orc_compare": (None, [D, dp, dp, N]),
            "orc_dot_prod": (D, [dp, dp, N]),
            "orc_max_norm": (D, [dp, N]),
            "orc_wsqr_sum": (D, [dp, dp, N]),
            "orc_wsqr_sum_mask": (D, [dp, dp, dp, N]),
            "orc_wrms_norm": (D, [dp, dp, N]),
            "orc_wrms_norm_mask": (D, [dp, dp, dp, N]),
            "orc_min": (D, [dp, N]),
            "orc_wl2_norm": (D, [dp, dp, N]),
            "orc_l1_norm": (D, [dp, N]),
            "orc_inv_test": (I, [dp, dp, N]),
            "orc_constr_mask": (I, [dp, dp, dp, N]),
            "orc_min_quotient": (D, [dp, dp, N]),
            "orc_linear_combination": (I, [I, dp, dpp, dp, N]),
            "orc_scale_add_multi": (I, [I, dp, dp, dpp, dpp, N]),
            "orc_dot_prod_multi": (I, [I, dp, dpp, dp, N]),
            "orc_linear_sum_vector_array": (I, [I, D, dpp, D, dpp, dpp, N]),
            "orc_scale_vector_array": (I, [I, dp, dpp, dpp, N]),
            "orc_const_vector_array": (I, [I, D, dpp, N]),
            "orc_wrms_norm_vector_array": (I, [I, dpp, dpp, dp, N]),
            "orc_wrms_norm_mask_vector_array": (I, [I, dpp, dpp, dp, dp, N]),
            "orc_scale_add_multi_vector_array": (I, [I, I, dp, dpp, dpp, dpp, N]),
            "orc_linear_combination_vector_array": (I, [I, I, dp, dpp, dpp, N]),
            "orc_mpi_wrms_from_local": (D, [dp, I, N]),
            "orc_fill_uniform": (None, [dp, N, C.c_uint32, D, D]),
            "orc_linear_sum_form": (I, [D, D, I, I]),
        }
        for k, (r, a) in sig.items():
            f = getattr(L, k)
            f.restype, f.argtypes = r, a

    # streaming
    def linear_sum(self, a, x, b, y, z): self.L.orc_linear_sum(a, _p(x), b, _p(y), _p(z), len(z))
    def const(self, c, z): self.L.orc_const(c, _p(z), len(z))
    def prod(self, x, y, z): self.L.orc_prod(_p(x), _p(y), _p(z), len(z))
    def div(self, x, y, z): self.L.orc_div(_p(x), _p(y), _p(z), len(z))
    def scale(self, c, x, z): self.L.orc_scale(c, _p(x), _p(z), len(z))
    def abs(self, x, z): self.L.orc_abs(_p(x), _p(z), len(z))
    def inv(self, x, z): self.L.orc_inv(_p(x), _p(z), len(z))
    def add_const(self, x, b, z): self.L.orc_add_const(_p(x), b, _p(z), len(z))
    def compare(self, c, x, z): self.L.orc_compare(c, _p(x), _p(z), len(z))
    # reductions
    def dot_prod(self, x, y): return self.L.orc_dot_prod(_p(x), _p(y), len(x))
    def max_norm(self, x): return self.L.orc_max_norm(_p(x), len(x))
    def wsqr_sum(self, x, w): return self.L.orc_wsqr_sum(_p(x), _p(w), len(x))
    def wsqr_sum_mask(self, x, w, id): return self.L.orc_wsqr_sum_mask(_p(x), _p(w), _p(id), len(x))
    def wrms_norm(self, x, w): return self.L.orc_wrms_norm(_p(x), _p(w), len(x))
    def wrms_norm_mask(self, x, w, id): return self.L.orc_wrms_norm_mask(_p(x), _p(w), _p(id), len(x))
    def min(self, x): return self.L.orc_min(_p(x), len(x))
    def wl2_norm(self, x, w): return self.L.orc_wl2_norm(_p(x), _p(w), len(x))
    def l1_norm(self, x): return self.L.orc_l1_norm(_p(x), len(x))
    def inv_test(self, x, z): return bool(self.L.orc_inv_test(_p(x), _p(z), len(x)))
    def constr_mask(self, c, x, m): return bool(self.L.orc_constr_mask(_p(c), _p(x), _p(m), len(x)))
    def min_quotient(self, num, den): return self.L.orc_min_quotient(_p(num), _p(den), len(num))

    # fused (lists of arrays; identity of list objects expresses array aliasing)
    def linear_combination(self, c, X, z):
        return self.L.orc_linear_combination(len(X), _coef(c), _tab(X), _p(z), len(z))

    def scale_add_multi(self, a, x, Y, Z):
        ty = _tab(Y)
        tz = ty if Z is Y else _tab(Z)
        return self.L.orc_scale_add_multi(len(Y), _coef(a), _p(x), ty, tz, len(x))

    def dot_prod_multi(self, x, Y):
        out = np.zeros(len(Y))
        self.L.orc_dot_prod_multi(len(Y), _p(x), _tab(Y), _p(out), len(x))
        return out

    def linear_sum_vector_array(self, a, X, b, Y, Z):
        tx, ty = _tab(X), _tab(Y)
        tz = tx if Z is X else ty if Z is Y else _tab(Z)
        return self.L.orc_linear_sum_vector_array(len(Z), a, tx, b, ty, tz, len(Z[0]))

    def scale_vector_array(self, c, X, Z):
        tx = _tab(X)
        tz = tx if Z is X else _tab(Z)
        return self.L.orc_scale_vector_array(len(Z), _coef(c), tx, tz, len(Z[0]))

    def const_vector_array(self, c, Z):
        return self.L.orc_const_vector_array(len(Z), c, _tab(Z), len(Z[0]))

    def wrms_norm_vector_array(self, X, W):
        out = np.zeros(len(X))
        self.L.orc_wrms_norm_vector_array(len(X), _tab(X), _tab(W), _p(out), len(X[0]))
        return out

    def wrms_norm_mask_vector_array(self, X, W, id):
        out = np.zeros(len(X))
        self.L.orc_wrms_norm_mask_vector_array(len(X), _tab(X), _tab(W), _p(id), _p(out), len(X[0]))
        return out

    def scale_add_multi_vector_array(self, a, X, Y, Z):
        nsum, nvec = len(Y), len(X)
        ty = _tab([Y[j][i] for j in range(nsum) for i in range(nvec)])
        tz = ty if (Z is Y or Z[0] is Y[0]) else _tab([Z[j][i] for j in range(nsum) for i in range(nvec)])
        return self.L.orc_scale_add_multi_vector_array(nvec, nsum, _coef(a), _tab(X), ty, tz, len(X[0]))

    def linear_combination_vector_array(self, c, X, Z):
        nsum, nvec = len(X), len(Z)
        flat = [X[i][j] for i in range(nsum) for j in range(nvec)]
        tx = _tab(flat)
        if X[0] is Z:
            tz = C.cast(tx, dpp)  # row 0 of the flattened table IS Z
        else:
            tz = _tab(Z)
        return self.L.orc_linear_combination_vector_array(nvec, nsum, _coef(c), tx, tz, len(Z[0]))

    def fill_uniform(self, n, seed, lo=-1.0, hi=1.0):
        x = np.empty(n)
        self.L.orc_fill_uniform(_p(x), n, seed, lo, hi)
        return x

    def mpi_wrms_from_local(self, sums, global_n):
        s = np.ascontiguousarray(sums, dtype=np.float64)
        return self.L.orc_mpi_wrms_from_local(_p(s), len(s), global_n)

    def linear_sum_form(self, a, b, z_is_x, z_is_y):
        return self.L.orc_linear_sum_form(a, b, int(z_is_x), int(z_is_y))


class RefSerial:
    """The real reference: N_V*_Serial on N_VMake_Serial-wrapped numpy arrays."""

    def __init__(self):
        if not REF_SO.exists():
            raise FileNotFoundError(REF_SO)
        L = C.CDLL(str(REF_SO), mode=C.RTLD_GLOBAL)
        self.L = L
        V = C.c_void_p
        D, I, N = C.c_double, C.c_int, C.c_int64
        Vp = C.POINTER(V)
        Vpp = C.POINTER(Vp)
        L.SUNContext_Create.restype, L.SUNContext_Create.argtypes = I, [I, C.POINTER(V)]
        L.N_VMake_Serial.restype, L.N_VMake_Serial.argtypes = V, [N, dp, V]
        L.N_VDestroy_Serial.restype, L.N_VDestroy_Serial.argtypes = None, [V]
        L.N_VEnableFusedOps_Serial.restype, L.N_VEnableFusedOps_Serial.argtypes = I, [V, I]
        self.ctx = V()
        assert L.SUNContext_Create(0, C.byref(self.ctx)) == 0
        sig = {
            "N_VLinearSum_Serial": (None, [D, V, D, V, V]),
            "N_VConst_Serial": (None, [D, V]),
            "N_VProd_Serial": (None, [V, V, V]),
            "N_VDiv_Serial": (None, [V, V, V]),
            "N_VScale_Serial": (None, [D, V, V]),
            "N_VAbs_Serial": (None, [V, V]),
            "N_VInv_Serial": (None, [V, V]),
            "N_VAddConst_Serial": (None, [V, D, V]),
            "N_VCompare_Serial": (None, [D, V, V]),
            "N_VDotProd_Serial": (D, [V, V]),
            "N_VMaxNorm_Serial": (D, [V]),
            "N_VWSqrSumLocal_Serial": (D, [V, V]),
            "N_VWSqrSumMaskLocal_Serial": (D, [V, V, V]),
            "N_VWrmsNorm_Serial": (D, [V, V]),
            "N_VWrmsNormMask_Serial": (D, [V, V, V]),
            "N_VMin_Serial": (D, [V]),
            "N_VWL2Norm_Serial": (D, [V, V]),
            "N_VL1Norm_Serial": (D, [V]),
            "N_VInvTest_Serial": (I, [V, V]),
            "N_VConstrMask_Serial": (I, [V, V, V]),
            "N_VMinQuotient_Serial": (D, [V, V]),
            "N_VLinearCombination_Serial": (I, [I, dp, Vp, V]),
            "N_VScaleAddMulti_Serial": (I, [I, dp, V, Vp, Vp]),
            "N_VDotProdMulti_Serial": (I, [I, V, Vp, dp]),
            "N_VLinearSumVectorArray_Serial": (I, [I, D, Vp, D, Vp, Vp]),
            "N_VScaleVectorArray_Serial": (I, [I, dp, Vp, Vp]),
            "N_VConstVectorArray_Serial": (I, [I, D, Vp]),
            "N_VWrmsNormVectorArray_Serial": (I, [I, Vp, Vp, dp]),
            "N_VWrmsNormMaskVectorArray_Serial": (I, [I, Vp, Vp, V, dp]),
            "N_VScaleAddMultiVectorArray_Serial": (I, [I, I, dp, Vp, Vpp, Vpp]),
            "N_VLinearCombinationVectorArray_Serial": (I, [I, I, dp, Vpp, Vp]),
        }
        for k, (r, a) in sig.items():
            f = getattr(L, k)
            f.restype, f.argtypes = r, a
        self._V, self._Vp = V, Vp

    # one N_Vector handle per distinct numpy array object within a call, so that
    # handle identity (z == x) mirrors array-object identity
    class _Handles:
        def __init__(self, ref):
            self.ref, self.map = ref, {}

        def h(self, a: np.ndarray):
            k = id(a)
            if k not in self.map:
                v = self.ref.L.N_VMake_Serial(len(a), _p(a), self.ref.ctx)
                self.map[k] = (v, a)
            return self.map[k][0]

        def arr(self, arrs, cache):
            k = id(arrs)
            if k not in cache:
                cache[k] = (self.ref._V * len(arrs))(*[self.h(a) for a in arrs])
            return cache[k]

        def close(self):
            for v, _ in self.map.values():
                self.ref.L.N_VDestroy_Serial(v)

    def _call(self, fn):
        H = RefSerial._Handles(self)
        try:
            return fn(H)
        finally:
            H.close()

    def linear_sum(self, a, x, b, y, z): self._call(lambda H: self.L.N_VLinearSum_Serial(a, H.h(x), b, H.h(y), H.h(z)))
    def const(self, c, z): self._call(lambda H: self.L.N_VConst_Serial(c, H.h(z)))
    def prod(self, x, y, z): self._call(lambda H: self.L.N_VProd_Serial(H.h(x), H.h(y), H.h(z)))
    def div(self, x, y, z): self._call(lambda H: self.L.N_VDiv_Serial(H.h(x), H.h(y), H.h(z)))
    def scale(self, c, x, z): self._call(lambda H: self.L.N_VScale_Serial(c, H.h(x), H.h(z)))
    def abs(self, x, z): self._call(lambda H: self.L.N_VAbs_Serial(H.h(x), H.h(z)))
    def inv(self, x, z): self._call(lambda H: self.L.N_VInv_Serial(H.h(x), H.h(z)))
    def add_const(self, x, b, z): self._call(lambda H: self.L.N_VAddConst_Serial(H.h(x), b, H.h(z)))
    def compare(self, c, x, z): self._call(lambda H: self.L.N_VCompare_Serial(c, H.h(x), H.h(z)))
    def dot_prod(self, x, y): return self._call(lambda H: self.L.N_VDotProd_Serial(H.h(x), H.h(y)))
    def max_norm(self, x): return self._call(lambda H: self.L.N_VMaxNorm_Serial(H.h(x)))
    def wsqr_sum(self, x, w): return self._call(lambda H: self.L.N_VWSqrSumLocal_Serial(H.h(x), H.h(w)))
    def wsqr_sum_mask(self, x, w, id): return self._call(lambda H: self.L.N_VWSqrSumMaskLocal_Serial(H.h(x), H.h(w), H.h(id)))
    def wrms_norm(self, x, w): return self._call(lambda H: self.L.N_VWrmsNorm_Serial(H.h(x), H.h(w)))
    def wrms_norm_mask(self, x, w, id): return self._call(lambda H: self.L.N_VWrmsNormMask_Serial(H.h(x), H.h(w), H.h(id)))
    def min(self, x): return self._call(lambda H: self.L.N_VMin_Serial(H.h(x)))
    def wl2_norm(self, x, w): return self._call(lambda H: self.L.N_VWL2Norm_Serial(H.h(x), H.h(w)))
    def l1_norm(self, x): return self._call(lambda H: self.L.N_VL1Norm_Serial(H.h(x)))
    def inv_test(self, x, z): return bool(self._call(lambda H: self.L.N_VInvTest_Serial(H.h(x), H.h(z))))
    def constr_mask(self, c, x, m): return bool(self._call(lambda H: self.L.N_VConstrMask_Serial(H.h(c), H.h(x), H.h(m))))
    def min_quotient(self, num, den): return self._call(lambda H: self.L.N_VMinQuotient_Serial(H.h(num), H.h(den)))

    def linear_combination(self, c, X, z):
        return self._call(lambda H: self.L.N_VLinearCombination_Serial(len(X), _coef(c), H.arr(X, {}), H.h(z)))

    def scale_add_multi(self, a, x, Y, Z):
        def f(H):
            cache = {}
            return self.L.N_VScaleAddMulti_Serial(len(Y), _coef(a), H.h(x), H.arr(Y, cache), H.arr(Z, cache))
        return self._call(f)

    def dot_prod_multi(self, x, Y):
        out = np.zeros(len(Y))
        self._call(lambda H: self.L.N_VDotProdMulti_Serial(len(Y), H.h(x), H.arr(Y, {}), _p(out)))
        return out

    def linear_sum_vector_array(self, a, X, b, Y, Z):
        def f(H):
            cache = {}
            return self.L.N_VLinearSumVectorArray_Serial(len(Z), a, H.arr(X, cache), b, H.arr(Y, cache), H.arr(Z, cache))
        return self._call(f)

    def scale_vector_array(self, c, X, Z):
        def f(H):
            cache = {}
            return self.L.N_VScaleVectorArray_Serial(len(Z), _coef(c), H.arr(X, cache), H.arr(Z, cache))
        return self._call(f)

    def const_vector_array(self, c, Z):
        return self._call(lambda H: self.L.N_VConstVectorArray_Serial(len(Z), c, H.arr(Z, {})))

    def wrms_norm_vector_array(self, X, W):
        out = np.zeros(len(X))
        self._call(lambda H: self.L.N_VWrmsNormVectorArray_Serial(len(X), H.arr(X, {}), H.arr(W, {}), _p(out)))
        return out

    def wrms_norm_mask_vector_array(self, X, W, id):
        out = np.zeros(len(X))
        self._call(lambda H: self.L.N_VWrmsNormMaskVectorArray_Serial(len(X), H.arr(X, {}), H.arr(W, {}), H.h(id), _p(out)))
        return out

    def _arr2d(self, H, rows, cache, cache2):
        k = id(rows)
        if k not in cache2:
            ptrs = [C.cast(H.arr(r, cache), self._Vp) for r in rows]
            cache2[k] = (self._Vp * len(rows))(*ptrs)
        return cache2[k]

    def scale_add_multi_vector_array(self, a, X, Y, Z):
        def f(H):
            cache, cache2 = {}, {}
            return self.L.N_VScaleAddMultiVectorArray_Serial(len(X), len(Y), _coef(a), H.arr(X, cache),
                                                             self._arr2d(H, Y, cache, cache2),
                                                             self._arr2d(H, Z, cache, cache2))
        return self._call(f)

    def linear_combination_vector_array(self, c, X, Z):
        def f(H):
            cache, cache2 = {}, {}
            return self.L.N_VLinearCombinationVectorArray_Serial(len(Z), len(X), _coef(c),
                                                                 self._arr2d(H, X, cache, cache2), H.arr(Z, cache))
        return self._call(f)


def have_ref() -> bool:
    return REF_SO.exists()
