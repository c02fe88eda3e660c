"""Python host-side mirror of the N_Vector operator interface over the B200 kernels.

Same names, argument order and semantics as the reference's generic N_V*
functions (include/sundials/sundials_nvector.h:215-330 of the reference), so the
parity tests read like test/unit_tests/nvector/test_nvector.c.  PyTorch is used
only as plumbing: it owns the device allocations (fp64 CUDA tensors) and the
stream; ALL arithmetic runs in libsundials_nvecb200.so (hand-written sm_100a
kernels) through the C ABI of include/b200vec.h.  There is no torch/numpy
fallback -- without a GPU and the built library every op raises.

Multi-GPU (one process per GPU): `Context.init_comm()` attaches an NCCL
communicator created in C (unique id broadcast through torch.distributed);
vectors made with `distributed=True` are the local block of a contiguous 1-D
partition (MPIPlusX pattern, src/nvector/mpiplusx/nvector_mpiplusx.c:30) and
their reductions allreduce over NVLink.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import B200VEC_MAX, B200VEC_MIN, B200VEC_SUM, B200VecError, check

SUN_BIG_REAL = 1.7976931348623157e308


class Context:
    """Execution context: device + stream + reduction workspace (+ communicator)."""

    def __init__(self, device: Optional[int] = None, stream: Optional[torch.cuda.Stream] = None):
        if not torch.cuda.is_available():
            raise B200VecError("sundials_b200 needs a CUDA device (no CPU fallback)")
        self.lib = _lib.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        h = C.c_void_p()
        sp = C.c_void_p(stream.cuda_stream) if stream is not None else C.c_void_p(
            torch.cuda.current_stream(self.device).cuda_stream)
        with torch.cuda.device(self.device):
            check(self.lib.b200vec_ctx_create(C.byref(h), self.device, sp), "ctx_create")
        self.h = h
        self.rank, self.size = 0, 1

    # -- tuning / bookkeeping
    def set_tuning(self, key: str, value: int) -> None:
        check(self.lib.b200vec_ctx_set_tuning(self.h, key.encode(), int(value)), f"set_tuning({key})")

    def get_tuning(self, key: str) -> int:
        return int(self.lib.b200vec_ctx_get_tuning(self.h, key.encode()))

    def launch_count(self) -> int:
        return int(self.lib.b200vec_ctx_launch_count(self.h))

    def sync(self) -> None:
        check(self.lib.b200vec_ctx_sync(self.h), "ctx_sync")

    def set_stream(self, stream: torch.cuda.Stream) -> None:
        check(self.lib.b200vec_ctx_set_stream(self.h, C.c_void_p(stream.cuda_stream)), "set_stream")

    # -- communicator
    def init_comm(self) -> None:
        """Attach an NCCL communicator spanning torch.distributed's world."""
        import torch.distributed as dist

        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        rank, size = dist.get_rank(), dist.get_world_size()
        idbuf = (C.c_ubyte * _lib.UNIQUE_ID_BYTES)()
        if rank == 0:
            check(self.lib.b200vec_comm_get_unique_id(idbuf), "comm_get_unique_id")
        t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            t = t.cuda(self.device)
        dist.broadcast(t, src=0)
        raw = bytes(t.cpu().tolist())
        idbuf = (C.c_ubyte * _lib.UNIQUE_ID_BYTES).from_buffer_copy(raw)
        check(self.lib.b200vec_comm_init(self.h, idbuf, rank, size), "comm_init")
        self.rank, self.size = rank, size

    def allreduce_slots(self, count: int, op: int) -> None:
        check(self.lib.b200vec_allreduce(self.h, count, op), "allreduce")

    def fetch(self, count: int) -> list[float]:
        out = (C.c_double * count)()
        check(self.lib.b200vec_result_fetch(self.h, count, out), "result_fetch")
        return list(out)

    def close(self) -> None:
        if self.h:
            self.lib.b200vec_ctx_release(self.h)
            self.h = C.c_void_p()

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass


_default_ctx: dict[int, Context] = {}


def default_context() -> Context:
    dev = torch.cuda.current_device() if torch.cuda.is_available() else -1
    if dev not in _default_ctx:
        _default_ctx[dev] = Context()
    return _default_ctx[dev]


class NVector:
    """A length-n fp64 vector resident in HBM (a torch CUDA tensor owns the bytes)."""

    __slots__ = ("data", "ctx", "global_length", "distributed")

    def __init__(self, data: torch.Tensor, ctx: Optional[Context] = None, distributed: bool = False,
                 global_length: Optional[int] = None):
        if data.dtype != torch.float64 or not data.is_cuda or not data.is_contiguous() or data.dim() != 1:
            raise B200VecError("NVector wraps a contiguous 1-D float64 CUDA tensor")
        self.data = data
        self.ctx = ctx or default_context()
        self.distributed = bool(distributed) and self.ctx.size > 1
        if global_length is None:
            global_length = data.numel()
            if self.distributed:
                g = C.c_int64(global_length)
                check(self.ctx.lib.b200vec_allreduce_i64_host(self.ctx.h, C.byref(g), B200VEC_SUM), "allreduce_i64")
                global_length = g.value
        self.global_length = int(global_length)

    @property
    def ptr(self) -> int:
        return self.data.data_ptr()

    def __len__(self) -> int:
        return self.data.numel()


# ---------------------------------------------------------------- constructors
def N_VNew(length: int, ctx: Optional[Context] = None, distributed: bool = False) -> NVector:
    ctx = ctx or default_context()
    return NVector(torch.empty(length, dtype=torch.float64, device=f"cuda:{ctx.device}"), ctx, distributed)


def N_VMake(t: torch.Tensor, ctx: Optional[Context] = None, distributed: bool = False) -> NVector:
    return NVector(t, ctx, distributed)


def N_VClone(w: NVector) -> NVector:
    return NVector(torch.empty_like(w.data), w.ctx, w.distributed, w.global_length)


def N_VGetLength(v: NVector) -> int:
    return v.global_length


def N_VGetLocalLength(v: NVector) -> int:
    return len(v)


# ------------------------------------------------------------------ helpers
def _table(vs: Sequence[NVector]):
    arr = (C.c_void_p * len(vs))(*[v.ptr for v in vs])
    return arr


def _coef(c: Sequence[float]):
    return (C.c_double * len(c))(*[float(x) for x in c])


def _L(v: NVector):
    return v.ctx.lib


def _reduce(v: NVector, launch, op: int) -> float:
    """launch(result_ptr) runs the local kernel; combine across ranks if needed."""
    if v.distributed:
        launch(None)
        v.ctx.allreduce_slots(1, op)
        return v.ctx.fetch(1)[0]
    r = C.c_double()
    launch(C.byref(r))
    return r.value


def _rsqrt(x: float) -> float:  # SUNRsqrt
    return 0.0 if x <= 0.0 else math.sqrt(x)


# ------------------------------------------------------------- streaming ops
def N_VLinearSum(a: float, x: NVector, b: float, y: NVector, z: NVector) -> None:
    check(_L(z).b200vec_linear_sum(z.ctx.h, a, x.ptr, b, y.ptr, z.ptr, len(z)), "N_VLinearSum")


def N_VConst(c: float, z: NVector) -> None:
    check(_L(z).b200vec_const(z.ctx.h, c, z.ptr, len(z)), "N_VConst")


def N_VProd(x: NVector, y: NVector, z: NVector) -> None:
    check(_L(z).b200vec_prod(z.ctx.h, x.ptr, y.ptr, z.ptr, len(z)), "N_VProd")


def N_VDiv(x: NVector, y: NVector, z: NVector) -> None:
    check(_L(z).b200vec_div(z.ctx.h, x.ptr, y.ptr, z.ptr, len(z)), "N_VDiv")


def N_VScale(c: float, x: NVector, z: NVector) -> None:
    check(_L(z).b200vec_scale(z.ctx.h, c, x.ptr, z.ptr, len(z)), "N_VScale")


def N_VAbs(x: NVector, z: NVector) -> None:
    check(_L(z).b200vec_abs(z.ctx.h, x.ptr, z.ptr, len(z)), "N_VAbs")


def N_VInv(x: NVector, z: NVector) -> None:
    check(_L(z).b200vec_inv(z.ctx.h, x.ptr, z.ptr, len(z)), "N_VInv")


def N_VAddConst(x: NVector, b: float, z: NVector) -> None:
    check(_L(z).b200vec_add_const(z.ctx.h, x.ptr, b, z.ptr, len(z)), "N_VAddConst")


def N_VCompare(c: float, x: NVector, z: NVector) -> None:
    check(_L(z).b200vec_compare(z.ctx.h, c, x.ptr, z.ptr, len(z)), "N_VCompare")


# ---------------------------------------------------------------- reductions
def N_VDotProdLocal(x: NVector, y: NVector) -> float:
    r = C.c_double()
    check(_L(x).b200vec_dot_prod(x.ctx.h, x.ptr, y.ptr, len(x), C.byref(r)), "N_VDotProdLocal")
    return r.value


def N_VDotProd(x: NVector, y: NVector) -> float:
    return _reduce(x, lambda r: check(_L(x).b200vec_dot_prod(x.ctx.h, x.ptr, y.ptr, len(x), r), "N_VDotProd"),
                   B200VEC_SUM)


def N_VMaxNorm(x: NVector) -> float:
    return _reduce(x, lambda r: check(_L(x).b200vec_max_norm(x.ctx.h, x.ptr, len(x), r), "N_VMaxNorm"), B200VEC_MAX)


def N_VMin(x: NVector) -> float:
    return _reduce(x, lambda r: check(_L(x).b200vec_min(x.ctx.h, x.ptr, len(x), r), "N_VMin"), B200VEC_MIN)


def N_VL1Norm(x: NVector) -> float:
    return _reduce(x, lambda r: check(_L(x).b200vec_l1_norm(x.ctx.h, x.ptr, len(x), r), "N_VL1Norm"), B200VEC_SUM)


def N_VWSqrSumLocal(x: NVector, w: NVector) -> float:
    r = C.c_double()
    check(_L(x).b200vec_wsqr_sum(x.ctx.h, x.ptr, w.ptr, len(x), C.byref(r)), "N_VWSqrSumLocal")
    return r.value


def N_VWSqrSumMaskLocal(x: NVector, w: NVector, id: NVector) -> float:
    r = C.c_double()
    check(_L(x).b200vec_wsqr_sum_mask(x.ctx.h, x.ptr, w.ptr, id.ptr, len(x), C.byref(r)), "N_VWSqrSumMaskLocal")
    return r.value


def _wsqr(x: NVector, w: NVector, id: Optional[NVector]) -> float:
    if id is None:
        return _reduce(x, lambda r: check(_L(x).b200vec_wsqr_sum(x.ctx.h, x.ptr, w.ptr, len(x), r), "wsqr_sum"),
                       B200VEC_SUM)
    return _reduce(x, lambda r: check(_L(x).b200vec_wsqr_sum_mask(x.ctx.h, x.ptr, w.ptr, id.ptr, len(x), r),
                                      "wsqr_sum_mask"), B200VEC_SUM)


def N_VWrmsNorm(x: NVector, w: NVector) -> float:
    return _rsqrt(_wsqr(x, w, None) / x.global_length)


def N_VWrmsNormMask(x: NVector, w: NVector, id: NVector) -> float:
    return _rsqrt(_wsqr(x, w, id) / x.global_length)


def N_VWL2Norm(x: NVector, w: NVector) -> float:
    return _rsqrt(_wsqr(x, w, None))


def N_VInvTest(x: NVector, z: NVector) -> bool:
    return _reduce(x, lambda r: check(_L(x).b200vec_inv_test(x.ctx.h, x.ptr, z.ptr, len(x), r), "N_VInvTest"),
                   B200VEC_MIN) > 0.5


def N_VConstrMask(c: NVector, x: NVector, m: NVector) -> bool:
    return _reduce(x, lambda r: check(_L(x).b200vec_constr_mask(x.ctx.h, c.ptr, x.ptr, m.ptr, len(x), r),
                                      "N_VConstrMask"), B200VEC_MIN) > 0.5


def N_VMinQuotient(num: NVector, denom: NVector) -> float:
    return _reduce(num, lambda r: check(_L(num).b200vec_min_quotient(num.ctx.h, num.ptr, denom.ptr, len(num), r),
                                        "N_VMinQuotient"), B200VEC_MIN)


# ------------------------------------------------------------------ fused ops
def N_VLinearCombination(c: Sequence[float], X: Sequence[NVector], z: NVector) -> None:
    check(_L(z).b200vec_linear_combination(z.ctx.h, len(X), _coef(c), _table(X), z.ptr, len(z)),
          "N_VLinearCombination")


def N_VLinearCombinationSqNorm(c: Sequence[float], X: Sequence[NVector], z: NVector) -> float:
    """z = sum c_i X_i and return z . z, one kernel (b200vec_linear_combination_sqnorm)."""
    lib = _L(z)
    if z.distributed:
        check(lib.b200vec_linear_combination_sqnorm(z.ctx.h, len(X), _coef(c), _table(X), z.ptr, len(z), None),
              "N_VLinearCombinationSqNorm")
        z.ctx.allreduce_slots(1, B200VEC_SUM)
        return z.ctx.fetch(1)[0]
    r = C.c_double()
    check(lib.b200vec_linear_combination_sqnorm(z.ctx.h, len(X), _coef(c), _table(X), z.ptr, len(z), C.byref(r)),
          "N_VLinearCombinationSqNorm")
    return r.value


def N_VScaleAddMulti(a: Sequence[float], x: NVector, Y: Sequence[NVector], Z: Sequence[NVector]) -> None:
    check(_L(x).b200vec_scale_add_multi(x.ctx.h, len(Y), _coef(a), x.ptr, _table(Y), _table(Z), len(x)),
          "N_VScaleAddMulti")


def N_VDotProdMulti(x: NVector, Y: Sequence[NVector]) -> list[float]:
    n = len(Y)
    if x.distributed:
        check(_L(x).b200vec_dot_prod_multi(x.ctx.h, n, x.ptr, _table(Y), len(x), None), "N_VDotProdMulti")
        x.ctx.allreduce_slots(n, B200VEC_SUM)
        return x.ctx.fetch(n)
    out = (C.c_double * n)()
    check(_L(x).b200vec_dot_prod_multi(x.ctx.h, n, x.ptr, _table(Y), len(x), out), "N_VDotProdMulti")
    return list(out)


# ----------------------------------------------------------- vector-array ops
def N_VLinearSumVectorArray(a: float, X: Sequence[NVector], b: float, Y: Sequence[NVector],
                            Z: Sequence[NVector]) -> None:
    check(_L(Z[0]).b200vec_linear_sum_vector_array(Z[0].ctx.h, len(Z), a, _table(X), b, _table(Y), _table(Z),
                                                   int(Z is X), int(Z is Y), len(Z[0])), "N_VLinearSumVectorArray")


def N_VScaleVectorArray(c: Sequence[float], X: Sequence[NVector], Z: Sequence[NVector]) -> None:
    check(_L(Z[0]).b200vec_scale_vector_array(Z[0].ctx.h, len(Z), _coef(c), _table(X), _table(Z), len(Z[0])),
          "N_VScaleVectorArray")


def N_VConstVectorArray(c: float, Z: Sequence[NVector]) -> None:
    check(_L(Z[0]).b200vec_const_vector_array(Z[0].ctx.h, len(Z), c, _table(Z), len(Z[0])), "N_VConstVectorArray")


def _wrms_va(X: Sequence[NVector], W: Sequence[NVector], id: Optional[NVector]) -> list[float]:
    n, x0 = len(X), X[0]
    idp = id.ptr if id is not None else None
    if x0.distributed:
        check(_L(x0).b200vec_wsqr_sum_vector_array(x0.ctx.h, n, _table(X), _table(W), idp, len(x0), None), "wrms_va")
        x0.ctx.allreduce_slots(n, B200VEC_SUM)
        s = x0.ctx.fetch(n)
    else:
        out = (C.c_double * n)()
        check(_L(x0).b200vec_wsqr_sum_vector_array(x0.ctx.h, n, _table(X), _table(W), idp, len(x0), out), "wrms_va")
        s = list(out)
    return [_rsqrt(v / x0.global_length) for v in s]


def N_VWrmsNormVectorArray(X: Sequence[NVector], W: Sequence[NVector]) -> list[float]:
    return _wrms_va(X, W, None)


def N_VWrmsNormMaskVectorArray(X: Sequence[NVector], W: Sequence[NVector], id: NVector) -> list[float]:
    return _wrms_va(X, W, id)


def N_VScaleAddMultiVectorArray(a: Sequence[float], X: Sequence[NVector], Y: Sequence[Sequence[NVector]],
                                Z: Sequence[Sequence[NVector]]) -> None:
    """Y[j][i], Z[j][i]: j = sum index (len(a)), i = vector index (len(X))."""
    nsum, nvec = len(Y), len(X)
    flatY = [Y[j][i] for j in range(nsum) for i in range(nvec)]
    flatZ = [Z[j][i] for j in range(nsum) for i in range(nvec)]
    y_is_z = int(Y is Z or Y[0] is Z[0])
    check(_L(X[0]).b200vec_scale_add_multi_vector_array(X[0].ctx.h, nvec, nsum, _coef(a), _table(X), _table(flatY),
                                                        _table(flatZ), y_is_z, len(X[0])),
          "N_VScaleAddMultiVectorArray")


def N_VLinearCombinationVectorArray(c: Sequence[float], X: Sequence[Sequence[NVector]], Z: Sequence[NVector]) -> None:
    """X[i][j]: i = term index (len(c)), j = vector index (len(Z))."""
    nsum, nvec = len(X), len(Z)
    flatX = [X[i][j] for i in range(nsum) for j in range(nvec)]
    check(_L(Z[0]).b200vec_linear_combination_vector_array(Z[0].ctx.h, nvec, nsum, _coef(c), _table(flatX), _table(Z),
                                                           int(X[0] is Z), len(Z[0])),
          "N_VLinearCombinationVectorArray")
