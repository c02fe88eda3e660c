"""Backend-independent parity cases for every N_Vector op on the hot path.

A *backend* is any object with the method names of tests/_oracle.Oracle
(Oracle, RefSerial, or the GPU adapter in tests/_b200_backend.py) operating IN
PLACE on numpy float64 arrays.  Each case builds seeded inputs, runs one op in
one aliasing / scalar configuration and returns {"name": value} where values are
arrays (compared bit-for-bit: streaming results) or floats tagged as reductions
(compared bit-for-bit between CPU backends and for n <= exact threshold on the
GPU; within the stated tolerance otherwise).

The configurations follow the reference's own unit tests
(test/unit_tests/nvector/test_nvector.c: Test_N_VLinearSum :561-1058 has the 9+
aliasing/scalar cases, LinearCombination :2155-2438, ScaleAddMulti :2440-2700,
vector-array tests :3000-5600) plus the serial special-case splits.
"""
from __future__ import annotations

import numpy as np

NVEC = 5   # vectors in fused / array ops (covers a partial batch of 4 + 1)
NSUM = 3


def rng_vec(rng, n, lo=-1.0, hi=1.0):
    return rng.uniform(lo, hi, n)


def _nz(rng, n):
    """values bounded away from zero (divisors)"""
    v = rng.uniform(0.5, 2.0, n)
    s = rng.integers(0, 2, n) * 2 - 1
    return v * s


SCALARS = [(1.0, 1.0), (1.0, -1.0), (-1.0, 1.0), (1.0, 0.37), (2.5, 1.0), (-1.0, 0.37), (1.7, -1.0),
           (0.75, 0.75), (1.25, -1.25), (0.3, -2.1), (0.0, 0.0), (0.0, 1.3)]


def linear_sum_cases():
    cases = []
    for (a, b) in SCALARS:
        for alias in ("none", "z=x", "z=y", "x=y", "all"):
            def run(B, n, seed, a=a, b=b, alias=alias):
                rng = np.random.default_rng(seed)
                x, y, z = rng_vec(rng, n), rng_vec(rng, n), rng_vec(rng, n)
                if alias == "z=x": z = x
                elif alias == "z=y": z = y
                elif alias == "x=y": y = x
                elif alias == "all": y = x; z = x
                B.linear_sum(a, x, b, y, z)
                return {"z": z.copy()}
            cases.append((f"linear_sum[a={a},b={b},{alias}]", run))
    return cases


def streaming_cases():
    cases = list(linear_sum_cases())

    def mk(name, fn):
        cases.append((name, fn))

    def c_const(B, n, seed):
        z = np.full(n, np.nan)
        B.const(-3.25, z)
        return {"z": z.copy()}
    mk("const", c_const)

    for alias in ("none", "z=x", "z=y"):
        def c_prod(B, n, seed, alias=alias):
            rng = np.random.default_rng(seed)
            x, y, z = rng_vec(rng, n), rng_vec(rng, n), np.zeros(n)
            if alias == "z=x": z = x
            if alias == "z=y": z = y
            B.prod(x, y, z)
            return {"z": z.copy()}
        mk(f"prod[{alias}]", c_prod)

        def c_div(B, n, seed, alias=alias):
            rng = np.random.default_rng(seed)
            x, y, z = rng_vec(rng, n), _nz(rng, n), np.zeros(n)
            if alias == "z=x": z = x
            if alias == "z=y": z = y
            B.div(x, y, z)
            return {"z": z.copy()}
        mk(f"div[{alias}]", c_div)

    for c in (1.0, -1.0, 0.0, 2.5, -0.3):
        for alias in ("none", "z=x"):
            def c_scale(B, n, seed, c=c, alias=alias):
                rng = np.random.default_rng(seed)
                x, z = rng_vec(rng, n), np.zeros(n)
                if alias == "z=x": z = x
                B.scale(c, x, z)
                return {"z": z.copy()}
            mk(f"scale[c={c},{alias}]", c_scale)

    for op in ("abs", "inv", "add_const", "compare"):
        for alias in ("none", "z=x"):
            def c_un(B, n, seed, op=op, alias=alias):
                rng = np.random.default_rng(seed)
                x = _nz(rng, n) if op == "inv" else rng_vec(rng, n)
                z = np.zeros(n)
                if alias == "z=x": z = x
                if op == "abs": B.abs(x, z)
                elif op == "inv": B.inv(x, z)
                elif op == "add_const": B.add_const(x, -0.7, z)
                else: B.compare(0.5, x, z)
                return {"z": z.copy()}
            mk(f"{op}[{alias}]", c_un)
    return cases


def reduction_cases():
    cases = []

    def mk(name, fn):
        cases.append((name, fn))

    def c_dot(B, n, seed):
        rng = np.random.default_rng(seed)
        x, y = rng_vec(rng, n), rng_vec(rng, n)
        return {"r:dot": B.dot_prod(x, y), "abs:dot": float(np.abs(x * y).sum())}
    mk("dot_prod", c_dot)

    def c_dot_self(B, n, seed):
        rng = np.random.default_rng(seed)
        x = rng_vec(rng, n)
        return {"r:dot": B.dot_prod(x, x), "abs:dot": float((x * x).sum())}
    mk("dot_prod[x=y]", c_dot_self)

    def c_maxnorm(B, n, seed):
        x = rng_vec(np.random.default_rng(seed), n)
        return {"e:max": B.max_norm(x)}
    mk("max_norm", c_maxnorm)

    def c_min(B, n, seed):
        x = rng_vec(np.random.default_rng(seed), n)
        return {"e:min": B.min(x)}
    mk("min", c_min)

    def c_l1(B, n, seed):
        x = rng_vec(np.random.default_rng(seed), n)
        return {"r:l1": B.l1_norm(x), "abs:l1": float(np.abs(x).sum())}
    mk("l1_norm", c_l1)

    def c_wsqr(B, n, seed):
        rng = np.random.default_rng(seed)
        x, w = rng_vec(rng, n), rng.uniform(0.5, 2.0, n)
        return {"r:wsqr": B.wsqr_sum(x, w), "r:wrms": B.wrms_norm(x, w), "r:wl2": B.wl2_norm(x, w)}
    mk("wsqr/wrms/wl2", c_wsqr)

    def c_wsqrmask(B, n, seed):
        rng = np.random.default_rng(seed)
        x, w = rng_vec(rng, n), rng.uniform(0.5, 2.0, n)
        id = rng.integers(0, 2, n).astype(np.float64)
        return {"r:wsqrmask": B.wsqr_sum_mask(x, w, id), "r:wrmsmask": B.wrms_norm_mask(x, w, id)}
    mk("wsqr_mask/wrms_mask", c_wsqrmask)

    for zeros in (False, True):
        def c_invtest(B, n, seed, zeros=zeros):
            rng = np.random.default_rng(seed)
            x = _nz(rng, n)
            if zeros and n > 0:
                x[rng.integers(0, n, max(1, n // 7))] = 0.0
            z = np.full(n, 7.0)  # untouched where x == 0
            ok = B.inv_test(x, z)
            return {"e:ok": float(ok), "z": z.copy()}
        mk(f"inv_test[zeros={zeros}]", c_invtest)

    for viol in (False, True):
        def c_constr(B, n, seed, viol=viol):
            rng = np.random.default_rng(seed)
            c = rng.integers(-2, 3, n).astype(np.float64)
            if viol:
                x = rng_vec(rng, n)
                if n > 3:
                    x[:3] = 0.0
            else:
                # satisfy every constraint strictly
                x = np.where(c == 0, rng_vec(rng, n), np.sign(c) * rng.uniform(0.1, 1.0, n))
            m = np.full(n, 9.0)
            ok = B.constr_mask(c, x, m)
            return {"e:ok": float(ok), "m": m.copy()}
        mk(f"constr_mask[viol={viol}]", c_constr)

    for mode in ("mixed", "allzero"):
        def c_minq(B, n, seed, mode=mode):
            rng = np.random.default_rng(seed)
            num = rng_vec(rng, n)
            den = np.zeros(n) if mode == "allzero" else np.where(rng.integers(0, 3, n) == 0, 0.0, _nz(rng, n))
            return {"e:minq": B.min_quotient(num, den)}
        mk(f"min_quotient[{mode}]", c_minq)
    return cases


def fused_cases():
    cases = []

    def mk(name, fn):
        cases.append((name, fn))

    # --- LinearCombination: nvec 1,2,3,NVEC, 9, 19 ; z aliasing X[0]; c0 == 1
    for nv in (1, 2, 3, NVEC, 9, 19):
        for alias in ("none", "z=X0", "z=X0,c0=1"):
            def c_lc(B, n, seed, nv=nv, alias=alias):
                rng = np.random.default_rng(seed)
                X = [rng_vec(rng, n) for _ in range(nv)]
                c = list(rng.uniform(-2, 2, nv))
                z = np.zeros(n)
                if alias.startswith("z=X0"): z = X[0]
                if alias.endswith("c0=1"): c[0] = 1.0
                rc = B.linear_combination(c, X, z)
                return {"rc": rc, "z": z.copy()}
            mk(f"linear_combination[nv={nv},{alias}]", c_lc)

    # --- ScaleAddMulti
    for nv in (1, 2, NVEC, 9, 18):
        for alias in ("none", "Y=Z"):
            def c_sam(B, n, seed, nv=nv, alias=alias):
                rng = np.random.default_rng(seed)
                x = rng_vec(rng, n)
                Y = [rng_vec(rng, n) for _ in range(nv)]
                Z = Y if alias == "Y=Z" else [np.zeros(n) for _ in range(nv)]
                a = list(rng.uniform(-2, 2, nv))
                rc = B.scale_add_multi(a, x, Y, Z)
                return {"rc": rc, **{f"z{j}": Z[j].copy() for j in range(nv)}}
            mk(f"scale_add_multi[nv={nv},{alias}]", c_sam)

    # --- DotProdMulti (Y may contain x itself: classical Gram-Schmidt does that)
    for nv in (1, 2, 3, 4, NVEC, 8, 9, 21):   # 2 / 3-4 / 5-8 outputs are separate kernel instantiations
        def c_dpm(B, n, seed, nv=nv):
            rng = np.random.default_rng(seed)
            x = rng_vec(rng, n)
            Y = [rng_vec(rng, n) for _ in range(nv)]
            Y[-1] = x
            d = B.dot_prod_multi(x, Y)
            return {"r:dots": np.asarray(d, dtype=np.float64).copy(),
                    "abs:dots": np.array([float(np.abs(x * y).sum()) for y in Y])}
        mk(f"dot_prod_multi[nv={nv}]", c_dpm)

    # x itself at other positions of Y (the kernel reads it once, through the shared operand), twice, and absent
    for (nv, pos) in ((2, (0,)), (3, (1,)), (4, (0, 3)), (6, (2,)), (10, (9,)), (4, ())):
        def c_dpms(B, n, seed, nv=nv, pos=pos):
            rng = np.random.default_rng(seed)
            x = rng_vec(rng, n)
            Y = [rng_vec(rng, n) for _ in range(nv)]
            for p in pos:
                Y[p] = x
            d = B.dot_prod_multi(x, Y)
            return {"r:dots": np.asarray(d, dtype=np.float64).copy(),
                    "abs:dots": np.array([float(np.abs(x * y).sum()) for y in Y])}
        mk(f"dot_prod_multi_self[nv={nv},pos={pos}]", c_dpms)
    return cases


def vector_array_cases():
    cases = []

    def mk(name, fn):
        cases.append((name, fn))

    for nv in (1, NVEC):
        for (a, b) in SCALARS:
            for alias in ("none", "Z=X", "Z=Y"):
                def c_lsva(B, n, seed, nv=nv, a=a, b=b, alias=alias):
                    rng = np.random.default_rng(seed)
                    X = [rng_vec(rng, n) for _ in range(nv)]
                    Y = [rng_vec(rng, n) for _ in range(nv)]
                    Z = X if alias == "Z=X" else Y if alias == "Z=Y" else [np.zeros(n) for _ in range(nv)]
                    rc = B.linear_sum_vector_array(a, X, b, Y, Z)
                    return {"rc": rc, **{f"z{j}": Z[j].copy() for j in range(nv)}}
                mk(f"linear_sum_va[nv={nv},a={a},b={b},{alias}]", c_lsva)

    for nv in (1, NVEC):
        for alias in ("none", "Z=X"):
            def c_sva(B, n, seed, nv=nv, alias=alias):
                rng = np.random.default_rng(seed)
                X = [rng_vec(rng, n) for _ in range(nv)]
                Z = X if alias == "Z=X" else [np.zeros(n) for _ in range(nv)]
                c = [1.0, -1.0, 0.5, 2.0, -0.25][:nv]
                rc = B.scale_vector_array(c, X, Z)
                return {"rc": rc, **{f"z{j}": Z[j].copy() for j in range(nv)}}
            mk(f"scale_va[nv={nv},{alias}]", c_sva)

        def c_cva(B, n, seed, nv=nv):
            Z = [np.full(n, np.nan) for _ in range(nv)]
            rc = B.const_vector_array(1.5, Z)
            return {"rc": rc, **{f"z{j}": Z[j].copy() for j in range(nv)}}
        mk(f"const_va[nv={nv}]", c_cva)

    for nv in (1, 2, 3, 4, NVEC, 11):
        def c_wva(B, n, seed, nv=nv):
            rng = np.random.default_rng(seed)
            X = [rng_vec(rng, n) for _ in range(nv)]
            W = [rng.uniform(0.5, 2.0, n) for _ in range(nv)]
            id = rng.integers(0, 2, n).astype(np.float64)
            return {"r:nrm": np.asarray(B.wrms_norm_vector_array(X, W)).copy(),
                    "r:nrmmask": np.asarray(B.wrms_norm_mask_vector_array(X, W, id)).copy()}
        mk(f"wrms_norm_va[nv={nv}]", c_wva)

    for (nv, ns) in ((1, 1), (1, NSUM), (NVEC, 1), (NVEC, NSUM), (3, 6)):
        for alias in ("none", "Y=Z"):
            def c_samva(B, n, seed, nv=nv, ns=ns, alias=alias):
                rng = np.random.default_rng(seed)
                X = [rng_vec(rng, n) for _ in range(nv)]
                Y = [[rng_vec(rng, n) for _ in range(nv)] for _ in range(ns)]
                Z = Y if alias == "Y=Z" else [[np.zeros(n) for _ in range(nv)] for _ in range(ns)]
                a = list(rng.uniform(-2, 2, ns))
                rc = B.scale_add_multi_vector_array(a, X, Y, Z)
                return {"rc": rc, **{f"z{j}_{i}": Z[j][i].copy() for j in range(ns) for i in range(nv)}}
            mk(f"scale_add_multi_va[nv={nv},ns={ns},{alias}]", c_samva)

    for (nv, ns) in ((1, 1), (1, 2), (1, NSUM + 1), (NVEC, 1), (NVEC, 2), (NVEC, NSUM), (2, 18)):
        for alias in ("none", "Z=X0", "Z=X0,c0=1"):
            def c_lcva(B, n, seed, nv=nv, ns=ns, alias=alias):
                rng = np.random.default_rng(seed)
                X = [[rng_vec(rng, n) for _ in range(nv)] for _ in range(ns)]
                Z = X[0] if alias.startswith("Z=X0") else [np.zeros(n) for _ in range(nv)]
                c = list(rng.uniform(-2, 2, ns))
                if alias.endswith("c0=1"): c[0] = 1.0
                rc = B.linear_combination_vector_array(c, X, Z)
                return {"rc": rc, **{f"z{j}": Z[j].copy() for j in range(nv)}}
            mk(f"linear_combination_va[nv={nv},ns={ns},{alias}]", c_lcva)
    return cases


def all_cases():
    return streaming_cases() + reduction_cases() + fused_cases() + vector_array_cases()


def compare(name, got: dict, want: dict, exact_reductions: bool, n: int, rtol=1e-13):
    """Compare a case's outputs.  Arrays and 'e:' scalars: bit-exact.  'r:' values
    (sums): bit-exact if exact_reductions else |got-want| <= rtol * sum|terms|
    (north_star tolerance 1e-13 relative; the sum of |terms| is the natural scale:
    the serial reference itself carries ~sqrt(n)*eps of rounding relative to it)."""
    assert got.keys() == want.keys(), (name, got.keys(), want.keys())
    for k in want:
        g, w = got[k], want[k]
        if k.startswith("abs:"):
            continue
        if k.startswith("r:") and not exact_reductions:
            g = np.atleast_1d(np.asarray(g, dtype=np.float64))
            w = np.atleast_1d(np.asarray(w, dtype=np.float64))
            # scale = sum of |terms| when the case supplies it (cancelling dots),
            # else the value itself (sums of non-negative terms, norms)
            scale = want.get("abs:" + k[2:], None)
            scale = np.abs(w) if scale is None else np.atleast_1d(np.asarray(scale, dtype=np.float64))
            tol = rtol * np.maximum(scale, 1e-300)
            assert np.all(np.abs(g - w) <= tol), (name, k, g, w, tol)
        else:
            ga = np.atleast_1d(np.asarray(g, dtype=np.float64))
            wa = np.atleast_1d(np.asarray(w, dtype=np.float64))
            assert ga.shape == wa.shape, (name, k)
            same = ga.view(np.uint64) == wa.view(np.uint64)
            assert np.all(same), (name, k, int((~same).sum()), "mismatching elements; first",
                                  int(np.argmax(~same)), ga[np.argmax(~same)], wa[np.argmax(~same)])
