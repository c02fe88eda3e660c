"""ctypes view of the N_Vector-level plugin API (include/nvector_b200.h).

This is the reference-facing boundary: the same `N_V*_B200` functions sit in the
SUNDIALS ops table (`v->ops->nvlinearsum == N_VLinearSum_B200`), so calling them
here is calling exactly what CVODE/ARKODE/IDA/KINSOL call.  Vectors are opaque
`N_Vector` handles (c_void_p).  Used by bench.py (value / e2e legs) and tests.

`Api` is deliberately generic: given a loaded library and a function-name
suffix it binds `N_V<Op><suffix>`; with the reference library and suffix ""
the very same suite drives the reference's generic dispatch (bench.py --impl
reference), so both arms execute an identical op list.
"""
from __future__ import annotations

import ctypes as C

V = C.c_void_p
D = C.c_double
I = C.c_int
L = C.c_int64
dp = C.POINTER(C.c_double)
Vp = C.POINTER(V)
Vpp = C.POINTER(Vp)

# op name -> (restype, argtypes) for the ops-table entries
OPS = {
    "LinearSum": (None, [D, V, D, V, V]),
    "Const": (None, [D, V]),
    "Prod": (None, [V, V, V]),
    "Div": (None, [V, V, V]),
    "Scale": (None, [D, V, V]),
    "Abs": (None, [V, V]),
    "Inv": (None, [V, V]),
    "AddConst": (None, [V, D, V]),
    "DotProd": (D, [V, V]),
    "MaxNorm": (D, [V]),
    "WrmsNorm": (D, [V, V]),
    "WrmsNormMask": (D, [V, V, V]),
    "Min": (D, [V]),
    "WL2Norm": (D, [V, V]),
    "L1Norm": (D, [V]),
    "Compare": (None, [D, V, V]),
    "InvTest": (I, [V, V]),
    "ConstrMask": (I, [V, V, V]),
    "MinQuotient": (D, [V, V]),
    "LinearCombination": (I, [I, dp, Vp, V]),
    "ScaleAddMulti": (I, [I, dp, V, Vp, Vp]),
    "DotProdMulti": (I, [I, V, Vp, dp]),
    "LinearSumVectorArray": (I, [I, D, Vp, D, Vp, Vp]),
    "ScaleVectorArray": (I, [I, dp, Vp, Vp]),
    "ConstVectorArray": (I, [I, D, Vp]),
    "WrmsNormVectorArray": (I, [I, Vp, Vp, dp]),
    "WrmsNormMaskVectorArray": (I, [I, Vp, Vp, V, dp]),
    "ScaleAddMultiVectorArray": (I, [I, I, dp, Vp, Vpp, Vpp]),
    "LinearCombinationVectorArray": (I, [I, I, dp, Vpp, Vp]),
    "DotProdLocal": (D, [V, V]),
    "MaxNormLocal": (D, [V]),
    "MinLocal": (D, [V]),
    "L1NormLocal": (D, [V]),
    "WSqrSumLocal": (D, [V, V]),
    "WSqrSumMaskLocal": (D, [V, V, V]),
    "InvTestLocal": (I, [V, V]),
    "ConstrMaskLocal": (I, [V, V, V]),
    "MinQuotientLocal": (D, [V, V]),
    "DotProdMultiLocal": (I, [I, V, Vp, dp]),
    "Clone": (V, [V]),
    "Destroy": (None, [V]),
    "GetLength": (L, [V]),
}


class Api:
    def __init__(self, lib: C.CDLL, suffix: str):
        self.lib, self.suffix = lib, suffix
        for op, (res, args) in OPS.items():
            fn = getattr(lib, f"N_V{op}{suffix}")
            fn.restype, fn.argtypes = res, args
            setattr(self, op, fn)

    @staticmethod
    def varray(handles):
        return (V * len(handles))(*handles)

    @staticmethod
    def varray2d(rows):
        keep = [Api.varray(r) for r in rows]
        arr = (Vp * len(rows))(*[C.cast(k, Vp) for k in keep])
        arr._keep = keep
        return arr

    @staticmethod
    def coefs(c):
        return (D * len(c))(*[float(x) for x in c])


class B200Plugin(Api):
    """N_V*_B200 from libsundials_nvecb200.so plus its constructors/accessors."""

    DEVICE, MANAGED, PINNED = 0, 1, 2

    def __init__(self):
        from . import _lib

        lib = _lib.load()
        super().__init__(lib, "_B200")
        f = lib.N_VNewWithCtx_B200
        f.restype, f.argtypes = V, [L, I, V, V]
        for name, res, args in (
            ("N_VGetHostArrayPointer_B200", dp, [V]),
            ("N_VGetDeviceArrayPointer_B200", V, [V]),
            ("N_VCopyToDevice_B200", None, [V]),
            ("N_VCopyFromDevice_B200", None, [V]),
            ("N_VEnableFusedOps_B200", I, [V, I]),
            ("N_VGetCtx_B200", V, [V]),
            ("N_VMakeDistributed_B200", I, [V, L]),
            ("N_VGetLocalLength_B200", L, [V]),
            ("N_VSetHostArrayPointer_B200", None, [dp, V]),
        ):
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args

    def new(self, n: int, ctx=None, kind: int = 0, fused: bool = True):
        v = self.lib.N_VNewWithCtx_B200(n, kind, ctx, None)
        if not v:
            from ._lib import B200VecError

            raise B200VecError("N_VNewWithCtx_B200 returned NULL: " + self.lib.b200vec_last_error().decode())
        if fused:
            self.lib.N_VEnableFusedOps_B200(v, 1)
        return v

    def host(self, v, n):
        import numpy as np

        p = self.lib.N_VGetHostArrayPointer_B200(v)
        return np.ctypeslib.as_array(p, shape=(n,))

    def to_device(self, v):
        self.lib.N_VCopyToDevice_B200(v)

    def from_device(self, v):
        self.lib.N_VCopyFromDevice_B200(v)

    def drop_host(self, v):
        """release the lazily created pinned host mirror of a device vector"""
        self.lib.N_VSetHostArrayPointer_B200(None, v)

    def ctx_of(self, v):
        return self.lib.N_VGetCtx_B200(v)
