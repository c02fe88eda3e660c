"""Second GPU backend for tests/_cases.py: the SHIPPED C host layer.

tests/_b200_backend.py drives sundials_b200/nvector.py, a Python view of the kernel-level C ABI.
This adapter instead goes through `N_V*_B200` -- the functions of nvector_b200.c that sit in the
SUNDIALS ops table and that CVODE/ARKODE/IDA/KINSOL call: handle-identity aliasing analysis,
`N_Vector*` array identity (Z == Y), global-length division, the fused-op gather tables.  Every
distinct numpy array becomes one N_Vector handle (device memory + pinned host mirror), so aliasing
between operands is preserved as HANDLE identity exactly as in a reference program.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from sundials_b200.plugin import B200Plugin, V, Vp


class PluginBackend:
    def __init__(self, ctx_handle=None):
        self.P = B200Plugin()
        self.ctx = ctx_handle  # raw b200vec_ctx (c_void_p) or None -> the default context

    class _Up:
        def __init__(self, be):
            self.be, self.map, self.arrays = be, {}, {}

        def v(self, a: np.ndarray):
            k = id(a)
            if k not in self.map:
                P = self.be.P
                h = P.new(len(a), self.be.ctx, P.DEVICE, fused=True)
                if len(a):
                    P.host(h, len(a))[...] = a
                    P.to_device(h)
                self.map[k] = (h, a)
            return self.map[k][0]

        def arr(self, L):
            """N_Vector* for a python list; the SAME ctypes array for the same list object"""
            k = id(L)
            if k not in self.arrays:
                self.arrays[k] = (V * len(L))(*[self.v(a) for a in L])
            return self.arrays[k]

        def arr2d(self, LL):
            k = id(LL)
            if k not in self.arrays:
                rows = [self.arr(L) for L in LL]
                self.arrays[k] = ((Vp * len(LL))(*[C.cast(r, Vp) for r in rows]), rows)
            return self.arrays[k][0]

        def down(self):
            P = self.be.P
            for h, a in self.map.values():
                if len(a):
                    P.from_device(h)
                    a[...] = P.host(h, len(a))
                P.Destroy(h)

    def _run(self, fn):
        U = PluginBackend._Up(self)
        try:
            return fn(U)
        finally:
            U.down()

    @staticmethod
    def _coefs(c):
        return (C.c_double * len(c))(*[float(x) for x in c])

    def linear_sum(self, a, x, b, y, z): self._run(lambda U: self.P.LinearSum(a, U.v(x), b, U.v(y), U.v(z)))
    def const(self, c, z): self._run(lambda U: self.P.Const(c, U.v(z)))
    def prod(self, x, y, z): self._run(lambda U: self.P.Prod(U.v(x), U.v(y), U.v(z)))
    def div(self, x, y, z): self._run(lambda U: self.P.Div(U.v(x), U.v(y), U.v(z)))
    def scale(self, c, x, z): self._run(lambda U: self.P.Scale(c, U.v(x), U.v(z)))
    def abs(self, x, z): self._run(lambda U: self.P.Abs(U.v(x), U.v(z)))
    def inv(self, x, z): self._run(lambda U: self.P.Inv(U.v(x), U.v(z)))
    def add_const(self, x, b, z): self._run(lambda U: self.P.AddConst(U.v(x), b, U.v(z)))
    def compare(self, c, x, z): self._run(lambda U: self.P.Compare(c, U.v(x), U.v(z)))
    def dot_prod(self, x, y): return self._run(lambda U: self.P.DotProd(U.v(x), U.v(y)))
    def max_norm(self, x): return self._run(lambda U: self.P.MaxNorm(U.v(x)))
    def wsqr_sum(self, x, w): return self._run(lambda U: self.P.WSqrSumLocal(U.v(x), U.v(w)))
    def wsqr_sum_mask(self, x, w, id): return self._run(lambda U: self.P.WSqrSumMaskLocal(U.v(x), U.v(w), U.v(id)))
    def wrms_norm(self, x, w): return self._run(lambda U: self.P.WrmsNorm(U.v(x), U.v(w)))
    def wrms_norm_mask(self, x, w, id): return self._run(lambda U: self.P.WrmsNormMask(U.v(x), U.v(w), U.v(id)))
    def min(self, x): return self._run(lambda U: self.P.Min(U.v(x)))
    def wl2_norm(self, x, w): return self._run(lambda U: self.P.WL2Norm(U.v(x), U.v(w)))
    def l1_norm(self, x): return self._run(lambda U: self.P.L1Norm(U.v(x)))
    def inv_test(self, x, z): return self._run(lambda U: bool(self.P.InvTest(U.v(x), U.v(z))))
    def constr_mask(self, c, x, m): return self._run(lambda U: bool(self.P.ConstrMask(U.v(c), U.v(x), U.v(m))))
    def min_quotient(self, num, den): return self._run(lambda U: self.P.MinQuotient(U.v(num), U.v(den)))

    def linear_combination(self, c, X, z):
        return self._run(lambda U: self.P.LinearCombination(len(X), self._coefs(c), U.arr(X), U.v(z)))

    def scale_add_multi(self, a, x, Y, Z):
        return self._run(lambda U: self.P.ScaleAddMulti(len(Y), self._coefs(a), U.v(x), U.arr(Y), U.arr(Z)))

    def dot_prod_multi(self, x, Y):
        def f(U):
            d = (C.c_double * len(Y))()
            assert self.P.DotProdMulti(len(Y), U.v(x), U.arr(Y), d) == 0
            return np.array(list(d))
        return self._run(f)

    def linear_sum_vector_array(self, a, X, b, Y, Z):
        return self._run(lambda U: self.P.LinearSumVectorArray(len(Z), a, U.arr(X), b, U.arr(Y), U.arr(Z)))

    def scale_vector_array(self, c, X, Z):
        return self._run(lambda U: self.P.ScaleVectorArray(len(Z), self._coefs(c), U.arr(X), U.arr(Z)))

    def const_vector_array(self, c, Z):
        return self._run(lambda U: self.P.ConstVectorArray(len(Z), c, U.arr(Z)))

    def wrms_norm_vector_array(self, X, W):
        def f(U):
            d = (C.c_double * len(X))()
            assert self.P.WrmsNormVectorArray(len(X), U.arr(X), U.arr(W), d) == 0
            return np.array(list(d))
        return self._run(f)

    def wrms_norm_mask_vector_array(self, X, W, id):
        def f(U):
            d = (C.c_double * len(X))()
            assert self.P.WrmsNormMaskVectorArray(len(X), U.arr(X), U.arr(W), U.v(id), d) == 0
            return np.array(list(d))
        return self._run(f)

    def scale_add_multi_vector_array(self, a, X, Y, Z):
        return self._run(lambda U: self.P.ScaleAddMultiVectorArray(len(X), len(Y), self._coefs(a), U.arr(X), U.arr2d(Y),
                                                                  U.arr2d(Z)))

    def linear_combination_vector_array(self, c, X, Z):
        return self._run(lambda U: self.P.LinearCombinationVectorArray(len(Z), len(X), self._coefs(c), U.arr2d(X),
                                                                      U.arr(Z)))
