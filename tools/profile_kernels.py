"""Launch each representative kernel a few times (target for ncu captures).
   ncu --set full --clock-control none --import-source on -k regex:'k_map|k_reduce|k_lincomb|k_scaleadd' \
       -s 14 -c 14 -o gpurun_out/prof python tools/profile_kernels.py --n 24   (second repetition)"""
import argparse
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from sundials_b200 import nvector as nv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=24)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
n = 1 << a.n
ctx = nv.default_context()
V = [nv.N_VMake(torch.rand(n, dtype=torch.float64, device="cuda") + 0.5, ctx) for _ in range(26)]
c = [0.3 + 0.1 * i for i in range(8)]
for _ in range(a.reps):
    nv.N_VLinearSum(0.3, V[0], -2.1, V[1], V[2])            # k_map<4,4,2,FGeneral>
    nv.N_VScale(2.5, V[3], V[4])                             # k_map<4,4,1,FScale>
    nv.N_VConst(1.5, V[5])                                   # k_map<4,4,0,FConst>
    nv.N_VDotProd(V[6], V[7])                                # k_reduce<4,4,RDot>
    nv.N_VMaxNorm(V[8])                                      # k_reduce<4,4,RMaxNorm>
    nv.N_VLinearCombination(c, V[:8], V[9])                  # k_lincomb_rows<4>
    nv.N_VScaleAddMulti(c, V[10], V[:8], V[11:19])           # k_scaleadd_rows<4>
    nv.N_VDotProdMulti(V[19], V[:8])                         # k_reduce_multi<4,0>
    nv.N_VWrmsNormVectorArray(V[:8], V[11:19])               # k_reduce_multi<4,1>
    nv.N_VLinearCombinationSqNorm([1.0] + c[:5], [V[20]] + V[:5], V[20])   # k_lincomb_sqnorm<4> (fused CGS step, k = 5)
    nv.N_VDotProdMulti(V[21], V[:20] + [V[21]])              # k_reduce_multi<4,0,24,1>: 21-wide (GMRES maxl = 20)
    r = C.c_double()
    ctx.lib.b200vec_axpy_dot(ctx.h, -0.37, V[22].ptr, V[23].ptr, V[24].ptr, n, C.byref(r))   # k_reduce<4,2,RAxpyDot>
    ctx.lib.b200vec_ewt_set(ctx.h, 1e-5, 1e-10, None, V[22].ptr, V[25].ptr, n, C.byref(r))   # k_reduce<4,2,REwt<false>>
    nv.N_VInvTest(V[22], V[25])                              # k_reduce<4,4,RInvTest>, one tile per CTA
torch.cuda.synchronize()
print("done")
