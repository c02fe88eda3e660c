"""Multi-GPU parity of the distributed vector (run under torchrun on N GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29544 tests/dist_parity_gpu.py

Every rank builds the SAME global vectors from a seed, keeps its contiguous
block on its GPU as an NVECTOR_B200 marked distributed (N_VMakeDistributed_B200,
NCCL communicator created in C from a broadcast unique id), runs every reducing
op through the plugin boundary and checks the GLOBAL result against the CPU
oracle on the whole vector; streaming ops are checked bit-exact on the block.
Mirrors test/unit_tests/nvector/mpiplusx/test_nvector_mpiplusx.c and the
global-length answers of Test_N_VDotProdMultiAllReduce (test_nvector.c:5663-5794).
"""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from _oracle import Oracle  # noqa: E402
from sundials_b200 import _lib  # noqa: E402
from sundials_b200.partition import block_range  # noqa: E402
from sundials_b200.plugin import B200Plugin  # noqa: E402


def main():
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    import datetime

    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lrank}"), timeout=datetime.timedelta(seconds=120))
    lib = _lib.load()
    P = B200Plugin()
    orc = Oracle()
    ctx = C.c_void_p()
    _lib.check(lib.b200vec_ctx_create(C.byref(ctx), lrank, None), "ctx_create")
    idbuf = (C.c_ubyte * _lib.UNIQUE_ID_BYTES)()
    if rank == 0:
        _lib.check(lib.b200vec_comm_get_unique_id(idbuf), "unique_id")
    t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, src=0)
    idbuf = (C.c_ubyte * _lib.UNIQUE_ID_BYTES).from_buffer_copy(bytes(t.cpu().tolist()))
    _lib.check(lib.b200vec_comm_init(ctx, idbuf, rank, world), "comm_init")
    assert lib.b200vec_comm_size(ctx) == world and lib.b200vec_comm_rank(ctx) == rank

    fails = 0
    # both transports: cross-rank fold inside the reduction kernel over NVLink peer
    # memory (default), then ncclAllReduce after the kernel ("p2p" = 0 on all ranks)
    cases = [(p2p, n) for p2p in (1, 0) for n in (1, 1000, 3_000_001)]
    for p2p, n in cases:
        _lib.check(lib.b200vec_ctx_set_tuning(ctx, b"p2p", p2p), "set_tuning(p2p)")
        transport = lib.b200vec_comm_transport(ctx).decode()
        if rank == 0:
            print(f"-- n={n} transport={transport}", flush=True)
        rng = np.random.default_rng(4242)          # identical global data on all ranks
        gx, gy = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
        gw = rng.uniform(0.5, 2.0, n)
        gid = rng.integers(0, 2, n).astype(np.float64)
        gc = rng.integers(-2, 3, n).astype(np.float64)
        a, b = block_range(n, rank, world)
        nl = b - a

        def mk(g):
            v = P.new(nl, ctx, P.DEVICE, fused=True)
            if nl > 0:      # n=1 on 2+ ranks: some ranks own an EMPTY block and must still take part
                P.host(v, nl)[...] = g[a:b]
                P.to_device(v)
            assert lib.N_VMakeDistributed_B200(v, -1) == 0      # global length by allreduce
            return v

        x, y, w, idv, cn = mk(gx), mk(gy), mk(gw), mk(gid), mk(gc)
        z = P.Clone(x)
        assert P.GetLength(x) == n and lib.N_VGetLocalLength_B200(x) == nl
        assert P.GetLength(z) == n                              # clones stay distributed

        def close(name, got, want, scale):
            nonlocal fails
            ok = abs(got - want) <= 1e-13 * max(scale, 1e-300)
            if not ok:
                fails += 1
                print(f"[rank {rank}] MISMATCH {name}: {got!r} vs {want!r}")

        close("DotProd", P.DotProd(x, y), orc.dot_prod(gx, gy), np.abs(gx * gy).sum())
        close("L1Norm", P.L1Norm(x), orc.l1_norm(gx), orc.l1_norm(gx))
        close("WrmsNorm", P.WrmsNorm(x, w), orc.wrms_norm(gx, gw), orc.wrms_norm(gx, gw))
        close("WrmsNormMask", P.WrmsNormMask(x, w, idv), orc.wrms_norm_mask(gx, gw, gid), 1.0)
        close("WL2Norm", P.WL2Norm(x, w), orc.wl2_norm(gx, gw), orc.wl2_norm(gx, gw))
        if P.MaxNorm(x) != orc.max_norm(gx) or P.Min(x) != orc.min(gx):
            fails += 1
            print(f"[rank {rank}] MISMATCH max/min")
        if P.MinQuotient(x, y) != orc.min_quotient(gx, gy):
            fails += 1
            print(f"[rank {rank}] MISMATCH minquotient")
        # flags: a zero / violation that lives on ONE rank must be seen by all
        gxz = gx.copy()
        gxz[n - 1] = 0.0
        xz = mk(gxz)
        if P.InvTest(xz, z) != 0 or P.InvTest(x, z) != 1:
            fails += 1
            print(f"[rank {rank}] MISMATCH invtest")
        zo = np.empty(n)
        want = orc.constr_mask(gc, gx, zo)
        if bool(P.ConstrMask(cn, x, z)) != want:
            fails += 1
            print(f"[rank {rank}] MISMATCH constrmask")
        if nl > 0:
            P.from_device(z)
            if not np.array_equal(P.host(z, nl), zo[a:b]):
                fails += 1
                print(f"[rank {rank}] MISMATCH constrmask mask block")
        # fused reductions: ONE nv-wide allreduce
        dots = (C.c_double * 3)()
        assert P.DotProdMulti(3, x, P.varray([y, w, x]), dots) == 0
        wantd = orc.dot_prod_multi(gx, [gy, gw, gx])
        for j in range(3):
            close(f"DotProdMulti[{j}]", dots[j], wantd[j], n)
        nrm = (C.c_double * 2)()
        assert P.WrmsNormVectorArray(2, P.varray([x, y]), P.varray([w, w]), nrm) == 0
        close("WrmsNormVA[0]", nrm[0], orc.wrms_norm(gx, gw), 1.0)
        close("WrmsNormVA[1]", nrm[1], orc.wrms_norm(gy, gw), 1.0)
        # local ops stay local
        close("DotProdLocal", P.DotProdLocal(x, y), orc.dot_prod(gx[a:b].copy(), gy[a:b].copy()), nl)
        # single-buffer form: local multi-dot then explicit allreduce
        assert P.DotProdMultiLocal(3, x, P.varray([y, w, x]), dots) == 0
        f = lib.N_VDotProdMultiAllReduce_B200
        f.restype, f.argtypes = C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_double)]
        assert f(3, x, dots) == 0
        for j in range(3):
            close(f"DotProdMultiAllReduce[{j}]", dots[j], wantd[j], n)
        # streaming on the block: bit exact, no communication
        P.LinearSum(0.3, x, -2.1, y, z)
        zo = np.empty(n)
        orc.linear_sum(0.3, gx, -2.1, gy, zo)
        if nl > 0:
            P.from_device(z)
            if not np.array_equal(P.host(z, nl).view(np.uint64), zo[a:b].view(np.uint64)):
                fails += 1
                print(f"[rank {rank}] MISMATCH linear_sum block")
        # fused Gram-Schmidt pieces on the distributed vectors, against the oracle on the WHOLE vectors:
        # (a) combination + squared norm, (b) update + next projection, (c) whole chained columns
        lib.N_VLinearCombinationSqNorm_B200.restype = C.c_int
        lib.N_VLinearCombinationSqNorm_B200.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_void_p),
                                                        C.c_void_p, C.POINTER(C.c_double)]
        lib.N_VAxpyDot_B200.restype = C.c_int
        lib.N_VAxpyDot_B200.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        lib.N_VModifiedGSSweep_B200.restype = C.c_int
        lib.N_VModifiedGSSweep_B200.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.c_void_p, C.POINTER(C.c_double),
                                                C.POINTER(C.c_double)]
        lib.N_VClassicalGSStep_B200.restype = C.c_int
        lib.N_VClassicalGSStep_B200.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                                C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        sq, dd = C.c_double(), C.c_double()
        gz = gx.copy()
        zz = mk(gz)
        cc = [1.0, -0.37, 0.61]
        assert lib.N_VLinearCombinationSqNorm_B200(3, P.coefs(cc), P.varray([zz, y, w]), zz, C.byref(sq)) == 0
        orc.linear_combination(cc, [gz, gy, gw], gz)
        close("LinearCombinationSqNorm", sq.value, orc.dot_prod(gz, gz), orc.dot_prod(gz, gz))
        assert lib.N_VAxpyDot_B200(-0.25, y, zz, w, C.byref(dd)) == 0
        orc.linear_sum(1.0, gz, -0.25, gy, gz)
        close("AxpyDot", dd.value, orc.dot_prod(gw, gz), np.abs(gw * gz).sum())
        # modified sweep against {y, w} and classical step against {y, w, zz}: mirror the reference's op sequence
        hcol, nrm2 = (C.c_double * 2)(), (C.c_double * 2)()
        assert lib.N_VModifiedGSSweep_B200(2, P.varray([y, w]), zz, hcol, nrm2) == 0
        want_n0 = orc.dot_prod(gz, gz)
        h0 = orc.dot_prod(gy, gz)
        orc.linear_sum(1.0, gz, -h0, gy, gz)
        h1 = orc.dot_prod(gw, gz)
        orc.linear_sum(1.0, gz, -h1, gw, gz)
        close("MGS sweep v.v", nrm2[0], want_n0, want_n0)
        close("MGS sweep h0", hcol[0], h0, np.abs(gy * gz).sum() + abs(h0))
        close("MGS sweep h1", hcol[1], h1, np.abs(gw * gz).sum() + abs(h1) + abs(h0) * np.abs(gw * gy).sum())
        close("MGS sweep norm", nrm2[1], orc.dot_prod(gz, gz), orc.dot_prod(gz, gz))
        if nl > 0:      # the vector itself stayed within rounding of the serial sequence
            P.from_device(zz)
            if not np.allclose(P.host(zz, nl), gz[a:b], rtol=0, atol=1e-12 * max(1.0, float(np.abs(gz).max()))):
                fails += 1
                print(f"[rank {rank}] MISMATCH modified Gram-Schmidt sweep vector block")
        # classical step on a fresh vector: dots against {y, w, itself}, then the combination + norm
        gq = gx.copy()
        q = mk(gq)
        d3 = (C.c_double * 3)()
        assert lib.N_VClassicalGSStep_B200(3, q, P.varray([y, w, q]), P.varray([q, y, w]), q, d3, C.byref(sq)) == 0
        wd = orc.dot_prod_multi(gq, [gy, gw, gq])
        for j, g2 in enumerate((gy, gw, gq)):
            close(f"CGS step dots[{j}]", d3[j], wd[j], np.abs(gq * g2).sum())
        orc.linear_combination([1.0, -wd[0], -wd[1]], [gq, gy, gw], gq)
        wn = orc.dot_prod(gq, gq)
        if abs(sq.value - wn) > 1e-10 * wn:   # the coefficients differ in their last bits: amplified by |d| |y| / |z|
            fails += 1
            print(f"[rank {rank}] MISMATCH CGS step norm: {sq.value!r} vs {wn!r}")
        if nl > 0:
            P.from_device(q)
            if not np.allclose(P.host(q, nl), gq[a:b], rtol=0, atol=1e-10 * max(1.0, float(np.abs(gq).max()))):
                fails += 1
                print(f"[rank {rank}] MISMATCH classical Gram-Schmidt step vector block")
        P.Destroy(q)
        P.Destroy(zz)
        # identical scalars on every rank (integrators must branch identically)
        v = torch.tensor([P.WrmsNorm(x, w)], dtype=torch.float64, device="cuda")
        lst = [torch.zeros_like(v) for _ in range(world)]
        dist.all_gather(lst, v)
        if not all(float(q) == float(lst[0]) for q in lst):
            fails += 1
            print(f"[rank {rank}] results differ between ranks")
        for h in (x, y, w, idv, cn, z, xz):
            P.Destroy(h)

    tf = torch.tensor([fails], device="cuda")
    dist.all_reduce(tf)
    if rank == 0:
        print("DIST PARITY", "OK" if int(tf.item()) == 0 else f"FAILED ({int(tf.item())})", f"world={world}")
    lib.b200vec_comm_finalize(ctx)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(tf.item()) == 0 else 1)


if __name__ == "__main__":
    main()
