"""Re-hosted benchmarks/advection_reaction_3D (apps/advection_reaction_3D: ARKODE ERK / DIRK /
IMEX-ARK and CVODE BDF / Adams on NVECTOR_B200, advection + reaction + upwind halo exchange in
one sm_100a kernel) against the REFERENCE.

Goldens (tests/golden/advection_reaction_3D, made here by tests/golden/make_ar3d_golden.py):
  rhs_*.npz    outputs of the reference's own SetIC / Advection / Reaction / AdvectionReaction /
               SolveReactionLinSys on seeded inputs                       -> kernels, BIT-EXACT
  *.out        stdout of the unmodified benchmark on the CPU              -> byte-identical table
               and statistics for <= 4096 unknowns (exact-order reductions)
  *.final.npz  its final solution (%.16e = exact doubles)                 -> BIT-EXACT state
Larger sizes: kernels bit-exact against the oracle (oracle/ar3d_oracle.py, pinned to the same
fixtures by tests/test_ar3d_oracle_cpu.py); full runs within the integrator's tolerance.
"""
import json
import os
import re
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "apps" / "advection_reaction_3D"))
GOLD = ROOT / "tests" / "golden" / "advection_reaction_3D"
MANIFEST = json.loads((GOLD / "MANIFEST.json").read_text())
RUN = ROOT / "apps" / "advection_reaction_3D" / "run.py"
RHS_CASES = sorted(p.stem for p in GOLD.glob("rhs_*.npz"))


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64).ravel()


def _same(a, b):
    return np.array_equal(_bits(a), _bits(b))


def _mismatch(a, b):
    a, b = np.asarray(a).ravel(), np.asarray(b).ravel()
    bad = np.nonzero(_bits(a) != _bits(b))[0]
    if bad.size == 0:
        return "identical"
    i = bad[0]
    return f"{bad.size} of {a.size} differ; first at {i}: {a[i]!r} vs {b[i]!r}"


@pytest.fixture(scope="module")
def ctx():
    import torch

    import run as ar

    torch.cuda.set_device(0)
    return ar.make_context(0, 0, 1)


def _dev(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).ravel()).cuda()


def _kernels_vs(ctx, n, c, gamma, y, b, want, **plan_kw):
    """all building blocks on one rank against `want` = dict(ic, fe, fi, f, x)"""
    import torch

    import run as ar

    plan = ar.Plan(ctx, npts=n, c=c, **plan_kw)
    try:
        assert plan.neq_loc == 3 * n ** 3
        yd, bd = _dev(y), _dev(b)
        out = torch.full_like(yd, -7.0)
        plan.set_ic(out)
        assert _same(out.cpu().numpy(), want["ic"]), "SetIC: " + _mismatch(out.cpu().numpy(), want["ic"])
        for which, key in ((ar.RHS_ADVECTION, "fe"), (ar.RHS_REACTION, "fi"), (ar.RHS_ADVECTION_REACTION, "f")):
            out.fill_(-7.0)
            plan.rhs(which, yd, out)
            got = out.cpu().numpy()
            assert _same(got, want[key]), f"{key}: " + _mismatch(got, want[key])
        out.fill_(-7.0)
        plan.psolve(yd, bd, out, gamma)
        assert _same(out.cpu().numpy(), want["x"]), "psolve: " + _mismatch(out.cpu().numpy(), want["x"])
        # in place (x aliases b), as the task-local Newton solver calls it
        bb = bd.clone()
        plan.psolve(yd, bb, bb, gamma)
        assert _same(bb.cpu().numpy(), want["x"]), "psolve in place"
        return plan.fast
    finally:
        plan.close()


@pytest.mark.parametrize("tag", RHS_CASES)
@pytest.mark.parametrize("generic", [0, 1])
def test_kernels_bit_exact_against_reference_functions(ctx, tag, generic):
    g = np.load(GOLD / f"{tag}.npz")
    n, c, gamma = int(g["npts"]), float(g["c"]), float(g["gamma"])
    fast = _kernels_vs(ctx, n, c, gamma, g["y"], g["b"], g, force_generic=generic)
    assert fast == (not generic and c > 0 and n % 4 == 0)


def _oracle_case(n, c, gamma, seed):
    import ar3d_oracle as orc

    rng = np.random.default_rng(seed)
    p = orc.params(c=c)
    d = p["xmax"] / n
    y = np.stack([1 + 0.2 * rng.uniform(-1, 1, (n, n, n)), 3.5 + 0.2 * rng.uniform(-1, 1, (n, n, n)),
                  3 + 0.2 * rng.uniform(-1, 1, (n, n, n))], axis=-1)
    b = rng.uniform(-1, 1, (n, n, n, 3))
    want = dict(ic=orc.initial_condition(n, p), fe=orc.advection(y, c, d, d, d), fi=orc.reaction(y, p),
                f=orc.advection_reaction(y, p, d, d, d), x=orc.solve_reaction_linsys(y, b, gamma, p))
    return y, b, want


@pytest.mark.parametrize("n,c,chunk", [(32, 0.01, 0), (32, 0.01, 1), (32, 0.01, 3), (32, 0.01, 64), (20, 2.0, 7),
                                       (64, 0.01, 0), (30, 0.01, 0), (16, -0.7, 0), (33, -0.01, 0), (4, 0.01, 0),
                                       (1, 0.01, 0), (2, -0.01, 0), (12, 0.0, 0)])
def test_kernels_bit_exact_against_oracle(ctx, n, c, chunk):
    # marching kernel (c > 0, npts % 4 == 0) at several planes-per-CTA incl. 1 and > nxl; generic
    # kernel otherwise; degenerate meshes (1, 2 points: every neighbour is the point itself / the
    # other point); c == 0 (advection identically zero)
    y, b, want = _oracle_case(n, c, 3e-3, 100 + n)
    _kernels_vs(ctx, n, c, 3e-3, y, b, want, planes_per_cta=chunk)


def test_component_masks(ctx):
    import torch

    import run as ar

    plan = ar.Plan(ctx, npts=8)
    try:
        for comp in range(3):
            m = torch.full((plan.neq_loc,), -1.0, dtype=torch.float64, device="cuda")
            plan.component_mask(comp, m)
            want = np.zeros((8 ** 3, 3))
            want[:, comp] = 1.0
            assert _same(m.cpu().numpy(), want)
    finally:
        plan.close()


def test_rhs_is_deterministic_and_linear_in_advection(ctx):
    # size-independent properties at a size the oracle does not visit in seconds (128^3, 6.3 M
    # unknowns): run-to-run identical bits; advection of a constant field is exactly zero; the
    # reaction of the steady state (u, v, w) = (A, B/A, B) is ~ 0
    import torch

    import run as ar

    n = 128
    plan = ar.Plan(ctx, npts=n)
    try:
        assert plan.fast
        g = torch.Generator(device="cuda").manual_seed(7)
        y = 1.0 + torch.rand(plan.neq_loc, dtype=torch.float64, device="cuda", generator=g)
        f1, f2 = torch.empty_like(y), torch.empty_like(y)
        plan.rhs(ar.RHS_ADVECTION_REACTION, y, f1)
        plan.rhs(ar.RHS_ADVECTION_REACTION, y, f2)
        assert torch.equal(f1.view(torch.int64), f2.view(torch.int64))
        fa, fr = torch.empty_like(y), torch.empty_like(y)
        plan.rhs(ar.RHS_ADVECTION, y, fa)
        plan.rhs(ar.RHS_REACTION, y, fr)
        # f = advection, then "+= reaction" (rhs3D.hpp:373): the fused kernel adds in that order
        assert torch.equal((fa + fr).view(torch.int64), f1.view(torch.int64))
        const = torch.full_like(y, 2.5)
        plan.rhs(ar.RHS_ADVECTION, const, fa)
        assert float(fa.abs().max()) == 0.0
        # generic kernel on the same input: identical bits
        gplan = ar.Plan(ctx, npts=n, force_generic=1)
        try:
            gplan.rhs(ar.RHS_ADVECTION_REACTION, y, f2)
            assert torch.equal(f1.view(torch.int64), f2.view(torch.int64))
        finally:
            gplan.close()
    finally:
        plan.close()


# ---------------------------------------------------------------- full runs
def _tail(text):
    i = text.index("          t   ")
    return text[i:]


def _run(args, extra=(), nproc=1, timeout=900, port=29571):
    if nproc == 1:
        cmd = [sys.executable, str(RUN), *args, *extra]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), str(RUN), *args, *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    return r.stdout


def _first_diff(a, b):
    for i, (x, y) in enumerate(zip(a.splitlines(), b.splitlines())):
        if x != y:
            return f"line {i + 1}:\n  b200     : {x}\n  reference: {y}"
    return f"length differs: {len(a)} vs {len(b)}"


EXACT = ["dirk_newton_8", "dirk_newton_nopre_8", "dirk_fixedpoint_8", "imex_newton_8", "imex_fixedpoint_8",
         "imex_newton_fused_8", "erk_8", "bdf_newton_8", "bdf_fixedpoint_8", "adams_8", "imex_newton_10_generic",
         "dirk_newton_cneg_8"]


@pytest.mark.parametrize("tag", EXACT)
def test_output_and_final_state_identical_to_reference_benchmark(tag):
    e = MANIFEST[tag]
    with tempfile.TemporaryDirectory() as td:
        out = _run(e["args"], ["--exact-threshold", "4096", "--save", "--output-dir", td])
        want = _tail((GOLD / f"{tag}.out").read_text())
        got = _tail(out)
        assert got == want, _first_diff(got, want)
        fin = np.load(GOLD / f"{tag}.final.npz")
        for s in "uvw":
            last = np.array((Path(td) / f"{s}.000000.txt").read_text().splitlines()[-1].split(), dtype=np.float64)
            assert _same(last, fin[s]), f"final {s}: " + _mismatch(last, fin[s])


def _table(text):
    rows = []
    for ln in _tail(text).splitlines():
        f = ln.split()
        if len(f) == 4 and re.fullmatch(r"[-+0-9.e]+", f[0]):
            rows.append([float(v) for v in f])
    return rows


def _stat(text, name):
    m = re.search(rf"{re.escape(name)}\s*=\s*(\d+)", text)
    assert m, name
    return int(m.group(1))


def _close_to(out, want, count_rel=0.05):
    tg, tw = _table(out), _table(want)
    assert len(tg) == len(tw) and len(tw) >= 2
    for g, w in zip(tg, tw):
        # the table prints 6 decimals; rtol 1e-6 integration
        assert abs(g[0] - w[0]) <= 1e-9
        for a, b in zip(g[1:], w[1:]):
            assert abs(a - b) <= 2e-5 * max(1.0, abs(b)), (g, w)
    for name in ("Internal solver steps", "Total number of nonlinear iterations"):
        g, w = _stat(out, name), _stat(want, name)
        assert abs(g - w) <= count_rel * w + 2, (name, g, w)


@pytest.mark.parametrize("tag", ["imex_newton_16", "dirk_newton_24_order4"])
def test_tree_reductions_stay_within_integrator_tolerance(tag):
    _close_to(_run(MANIFEST[tag]["args"]), (GOLD / f"{tag}.out").read_text())


def test_task_local_newton_agrees_with_global_newton():
    # the reference's own tl-newton cannot run in a non-MPI build (see make_ar3d_golden.py); both
    # solvers converge the same stage equations to the integrator's tolerance
    args = MANIFEST["imex_newton_8"]["args"]
    out = _run(args, ["--nls", "tl-newton", "--exact-threshold", "4096"])
    want = (GOLD / "imex_newton_8.out").read_text()
    tg, tw = _table(out), _table(want)
    assert len(tg) == len(tw)
    for g, w in zip(tg, tw):
        for a, b in zip(g[1:], w[1:]):
            assert abs(a - b) <= 5e-5 * max(1.0, abs(b)), (g, w)
    assert "Total number of linear iterations" not in _tail(out)


def _ngpus():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("nproc", [2, 4])
def test_n_ranks_kernels_bit_exact_and_runs_match(nproc):
    """2 ranks: the periodic west and east neighbour are the same GPU; 4 ranks: they differ."""
    if _ngpus() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
                        "--master-addr", "127.0.0.1", "--master-port", "29573", str(ROOT / "tests" / "ar3d_dist_gpu.py")],
                       capture_output=True, text=True, timeout=900, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "AR3D DIST OK" in r.stdout
    for tag in ("imex_newton_8", "dirk_newton_8", "imex_newton_16", "dirk_newton_cneg_8"):
        _close_to(_run(MANIFEST[tag]["args"], nproc=nproc, port=29575), (GOLD / f"{tag}.out").read_text(), count_rel=0.1)
