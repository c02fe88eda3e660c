"""N_VMakeManaged_B200 -- wrapping a user's cudaMallocManaged array (counterpart of
N_VMakeManaged_Cuda, nvector_cuda.cu:387).  Runs last in the suite (file name)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def be():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    from _b200_backend import B200Backend

    return B200Backend()


def test_make_managed_wraps_a_user_array(be):
    """N_VMakeManaged_B200 (counterpart of N_VMakeManaged_Cuda, nvector_cuda.cu:387): ONE
    cudaMallocManaged array per vector, not owned, host-coherent -- the host reads the result right
    after the op returns, and the array outlives the vector."""
    import ctypes as C

    from _oracle import REF_SO
    from sundials_b200 import _lib

    lib = _lib.load()
    core = C.CDLL(str(REF_SO), mode=C.RTLD_GLOBAL)
    core.SUNContext_Create.restype, core.SUNContext_Create.argtypes = C.c_int, [C.c_int, C.POINTER(C.c_void_p)]
    sctx = C.c_void_p()
    assert core.SUNContext_Create(0, C.byref(sctx)) == 0
    V, D = C.c_void_p, C.c_double
    for name, res, args in (("N_VMakeManaged_B200", V, [C.c_int64, V, V]), ("N_VIsManagedMemory_B200", C.c_int, [V]),
                            ("N_VLinearSum_B200", None, [D, V, D, V, V]), ("N_VDotProd_B200", D, [V, V]),
                            ("N_VGetLength_B200", C.c_int64, [V]), ("N_VDestroy_B200", None, [V])):
        f = getattr(lib, name)
        f.restype, f.argtypes = res, args
    n = 1000
    p = C.c_void_p()
    _lib.check(lib.b200vec_malloc_managed(be.ctx.h, 3 * n * 8, C.byref(p)), "malloc_managed")
    try:
        arr = np.ctypeslib.as_array((D * (3 * n)).from_address(p.value))
        rng = np.random.default_rng(5)
        arr[:n] = rng.uniform(-1, 1, n)
        arr[n:2 * n] = rng.uniform(-1, 1, n)
        arr[2 * n:] = 0.0
        x0, y0 = arr[:n].copy(), arr[n:2 * n].copy()
        vx, vy, vz = (lib.N_VMakeManaged_B200(n, V(p.value + k * n * 8), sctx) for k in range(3))
        assert vx and vy and vz
        assert lib.N_VIsManagedMemory_B200(vx) == 1 and lib.N_VGetLength_B200(vz) == n
        lib.N_VLinearSum_B200(2.0, vx, -1.0, vy, vz)            # host-coherent: no explicit sync or copy
        want = 2.0 * x0 - y0
        assert np.array_equal(arr[2 * n:].view(np.uint64), want.view(np.uint64))
        d = lib.N_VDotProd_B200(vx, vy)
        assert abs(d - float(np.dot(x0, y0))) <= 1e-13 * float(np.abs(x0 * y0).sum())
        for v in (vx, vy, vz):
            lib.N_VDestroy_B200(v)
        assert np.array_equal(arr[:n], x0) and np.array_equal(arr[2 * n:], want)   # not owned: still there
    finally:
        _lib.check(lib.b200vec_free_managed(be.ctx.h, p), "free_managed")
