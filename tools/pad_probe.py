"""HBM address-aliasing probe: time k_map-type ops and the diffusion RHS with the operands' base addresses
staggered by different paddings (all 256-byte aligned)."""
import ctypes as C, sys, os, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/apps/diffusion_2D')
from sundials_b200 import nvector as nv
import run as app
n = 1 << 24
ctx = nv.default_context()
def timed(fn, reps=30):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for pad_elems in (0, 32, 4096 // 8, 65536 // 8 + 32, (1 << 20) // 8 + 512, (1 << 21) // 8 + 96):
    K = 8
    big = torch.rand(K * (n + pad_elems) + 64, dtype=torch.float64, device='cuda')
    vs = [nv.N_VMake(big[i * (n + pad_elems): i * (n + pad_elems) + n], ctx) for i in range(K)]
    t_scale = timed(lambda i: nv.N_VScale(2.5, vs[i % 4], vs[4 + i % 4]))
    t_ls = timed(lambda i: nv.N_VLinearSum(0.3, vs[i % 3], -2.1, vs[3 + i % 2], vs[5 + i % 3]))
    t_dot = timed(lambda i: ctx.lib.b200vec_dot_prod(ctx.h, vs[i % 4].ptr, vs[4 + i % 4].ptr, n, None))
    print(f"pad {pad_elems*8:8d} B: Scale {t_scale:6.2f} us {16*n/t_scale/1e3:6.0f} GB/s | LinearSum {t_ls:6.2f} us {24*n/t_ls/1e3:6.0f} GB/s | Dot(kernel) {t_dot:6.2f} us {16*n/t_dot/1e3:6.0f} GB/s", flush=True)
    del vs, big
# diffusion RHS 8192^2 with u / f staggered
lib = app.load()
dctx = app.make_context(0, 0, 1)
lib.b200_diffusion2d_plan_create.restype = C.c_int
lib.b200_diffusion2d_plan_create.argtypes = [C.c_void_p, C.POINTER(app.Opts), C.POINTER(C.c_void_p)]
lib.b200_diffusion2d_rhs.restype = C.c_int
lib.b200_diffusion2d_rhs.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
N = 8192 * 8192
for pad_elems in (0, 4096 // 8, 65536 // 8 + 32, (1 << 20) // 8 + 512, (1 << 21) // 8 + 96):
    big = torch.rand(4 * (N + pad_elems) + 64, dtype=torch.float64, device='cuda')
    bufs = [big[i * (N + pad_elems): i * (N + pad_elems) + N] for i in range(4)]
    o = app.Opts(); lib.b200_diffusion2d_default_opts(C.byref(o)); o.nx = o.ny = 8192
    plan = C.c_void_p(); assert lib.b200_diffusion2d_plan_create(dctx, C.byref(o), C.byref(plan)) == 0
    t = timed(lambda i: lib.b200_diffusion2d_rhs(plan, 0.3, bufs[i % 2].data_ptr(), bufs[2 + i % 2].data_ptr()), 20)
    print(f"pad {pad_elems*8:8d} B: diffusion RHS {t:7.2f} us {16*N/t/1e3:6.0f} GB/s", flush=True)
    del bufs, big
