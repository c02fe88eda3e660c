"""GPU backend adapter for tests/_cases.py: same method names as the CPU
oracles, but every op runs in libsundials_nvecb200.so on cuda:0.

Per call: each distinct numpy array object is uploaded once to a fresh device
tensor (so aliasing between operands is preserved as pointer identity on the
device), the op runs through sundials_b200.nvector (C ABI), and every array is
downloaded back in place.  `misalign` (in doubles) offsets the device arrays to
exercise the 128-bit / 64-bit load paths.
"""
from __future__ import annotations

import numpy as np
import torch

from sundials_b200 import nvector as nv


class B200Backend:
    def __init__(self, ctx=None, misalign: int = 0):
        self.ctx = ctx or nv.default_context()
        self.misalign = misalign

    class _Up:
        def __init__(self, be):
            self.be, self.map = be, {}

        def v(self, a: np.ndarray) -> nv.NVector:
            k = id(a)
            if k not in self.map:
                off = self.be.misalign
                buf = torch.empty(len(a) + off + 4, dtype=torch.float64, device=f"cuda:{self.be.ctx.device}")
                t = buf[off:off + len(a)]
                t.copy_(torch.from_numpy(a))
                self.map[k] = (nv.NVector(t, self.be.ctx), a, buf)
            return self.map[k][0]

        def vs(self, arrs):
            return [self.v(a) for a in arrs]

        def down(self):
            self.be.ctx.sync()
            for vec, a, _ in self.map.values():
                a[...] = vec.data.cpu().numpy()

    def _run(self, fn):
        U = B200Backend._Up(self)
        r = fn(U)
        U.down()
        return r

    # identity-preserving list mapping (Z is Y  ->  same python list on the device side)
    @staticmethod
    def _lists(U, *lists):
        cache, out = {}, []
        for L in lists:
            if id(L) not in cache:
                cache[id(L)] = U.vs(L)
            out.append(cache[id(L)])
        return out

    @staticmethod
    def _lists2d(U, *lol):
        cache, cache2, out = {}, {}, []
        for LL in lol:
            if id(LL) not in cache2:
                rows = []
                for L in LL:
                    if id(L) not in cache:
                        cache[id(L)] = U.vs(L)
                    rows.append(cache[id(L)])
                cache2[id(LL)] = rows
            out.append(cache2[id(LL)])
        return out, cache

    def linear_sum(self, a, x, b, y, z): self._run(lambda U: nv.N_VLinearSum(a, U.v(x), b, U.v(y), U.v(z)))
    def const(self, c, z): self._run(lambda U: nv.N_VConst(c, U.v(z)))
    def prod(self, x, y, z): self._run(lambda U: nv.N_VProd(U.v(x), U.v(y), U.v(z)))
    def div(self, x, y, z): self._run(lambda U: nv.N_VDiv(U.v(x), U.v(y), U.v(z)))
    def scale(self, c, x, z): self._run(lambda U: nv.N_VScale(c, U.v(x), U.v(z)))
    def abs(self, x, z): self._run(lambda U: nv.N_VAbs(U.v(x), U.v(z)))
    def inv(self, x, z): self._run(lambda U: nv.N_VInv(U.v(x), U.v(z)))
    def add_const(self, x, b, z): self._run(lambda U: nv.N_VAddConst(U.v(x), b, U.v(z)))
    def compare(self, c, x, z): self._run(lambda U: nv.N_VCompare(c, U.v(x), U.v(z)))
    def dot_prod(self, x, y): return self._run(lambda U: nv.N_VDotProd(U.v(x), U.v(y)))
    def max_norm(self, x): return self._run(lambda U: nv.N_VMaxNorm(U.v(x)))
    def wsqr_sum(self, x, w): return self._run(lambda U: nv.N_VWSqrSumLocal(U.v(x), U.v(w)))
    def wsqr_sum_mask(self, x, w, id): return self._run(lambda U: nv.N_VWSqrSumMaskLocal(U.v(x), U.v(w), U.v(id)))
    def wrms_norm(self, x, w): return self._run(lambda U: nv.N_VWrmsNorm(U.v(x), U.v(w)))
    def wrms_norm_mask(self, x, w, id): return self._run(lambda U: nv.N_VWrmsNormMask(U.v(x), U.v(w), U.v(id)))
    def min(self, x): return self._run(lambda U: nv.N_VMin(U.v(x)))
    def wl2_norm(self, x, w): return self._run(lambda U: nv.N_VWL2Norm(U.v(x), U.v(w)))
    def l1_norm(self, x): return self._run(lambda U: nv.N_VL1Norm(U.v(x)))
    def inv_test(self, x, z): return self._run(lambda U: nv.N_VInvTest(U.v(x), U.v(z)))
    def constr_mask(self, c, x, m): return self._run(lambda U: nv.N_VConstrMask(U.v(c), U.v(x), U.v(m)))
    def min_quotient(self, num, den): return self._run(lambda U: nv.N_VMinQuotient(U.v(num), U.v(den)))

    def linear_combination(self, c, X, z):
        self._run(lambda U: nv.N_VLinearCombination(c, U.vs(X), U.v(z)))
        return 0

    def scale_add_multi(self, a, x, Y, Z):
        def f(U):
            dY, dZ = self._lists(U, Y, Z)
            nv.N_VScaleAddMulti(a, U.v(x), dY, dZ)
        self._run(f)
        return 0

    def dot_prod_multi(self, x, Y):
        return np.array(self._run(lambda U: nv.N_VDotProdMulti(U.v(x), U.vs(Y))))

    def linear_sum_vector_array(self, a, X, b, Y, Z):
        def f(U):
            dX, dY, dZ = self._lists(U, X, Y, Z)
            nv.N_VLinearSumVectorArray(a, dX, b, dY, dZ)
        self._run(f)
        return 0

    def scale_vector_array(self, c, X, Z):
        def f(U):
            dX, dZ = self._lists(U, X, Z)
            nv.N_VScaleVectorArray(c, dX, dZ)
        self._run(f)
        return 0

    def const_vector_array(self, c, Z):
        self._run(lambda U: nv.N_VConstVectorArray(c, U.vs(Z)))
        return 0

    def wrms_norm_vector_array(self, X, W):
        return np.array(self._run(lambda U: nv.N_VWrmsNormVectorArray(U.vs(X), U.vs(W))))

    def wrms_norm_mask_vector_array(self, X, W, id):
        return np.array(self._run(lambda U: nv.N_VWrmsNormMaskVectorArray(U.vs(X), U.vs(W), U.v(id))))

    def scale_add_multi_vector_array(self, a, X, Y, Z):
        def f(U):
            (dY, dZ), cache = self._lists2d(U, Y, Z)
            dX = cache.get(id(X)) or U.vs(X)
            nv.N_VScaleAddMultiVectorArray(a, dX, dY, dZ)
        self._run(f)
        return 0

    def linear_combination_vector_array(self, c, X, Z):
        def f(U):
            (dX,), cache = self._lists2d(U, X)
            dZ = cache.get(id(Z)) or U.vs(Z)
            nv.N_VLinearCombinationVectorArray(c, dX, dZ)
        self._run(f)
        return 0
