"""The CPU restatement of the advection_reaction_3D right-hand side (oracle/ar3d_oracle.py)
against outputs of the REFERENCE's own functions (tests/golden/advection_reaction_3D/rhs_*.npz,
made here by tests/c/ar3d_rhs_dump.cpp -> rhs3D.hpp by path): bit-for-bit."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
import ar3d_oracle as orc  # noqa: E402

GOLD = ROOT / "tests" / "golden" / "advection_reaction_3D"
CASES = sorted(p.stem for p in GOLD.glob("rhs_*.npz"))


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def _same(a, b):
    # +0.0 == -0.0 is not enough: compare the bit patterns
    return np.array_equal(_bits(a).ravel(), _bits(b).ravel())


def test_golden_fixtures_present():
    assert len(CASES) >= 4


@pytest.mark.parametrize("tag", CASES)
def test_oracle_matches_reference_functions(tag):
    g = np.load(GOLD / f"{tag}.npz")
    n, c, gamma = int(g["npts"]), float(g["c"]), float(g["gamma"])
    p = orc.params(c=c)
    d = p["xmax"] / n
    y = g["y"].reshape(n, n, n, 3)
    b = g["b"].reshape(n, n, n, 3)
    assert _same(orc.initial_condition(n, p), g["ic"]), "SetIC"
    assert _same(orc.advection(y, c, d, d, d), g["fe"]), "Advection"
    assert _same(orc.reaction(y, p), g["fi"]), "Reaction"
    assert _same(orc.advection_reaction(y, p, d, d, d), g["f"]), "AdvectionReaction"
    assert _same(orc.solve_reaction_linsys(y, b, gamma, p), g["x"]), "SolveReactionLinSys"


def test_slab_halo_form_equals_periodic_single_rank():
    # two slabs in x with the upstream neighbour's plane as halo: every point keeps its
    # arithmetic, but the slab's own first plane takes the FACE summation order -- so the
    # result differs from the one-rank run only there, and only in rounding
    rng = np.random.default_rng(3)
    n = 8
    y = rng.uniform(0.5, 1.5, (n, n, n, 3))
    d = 1.0 / n
    for c in (0.3, -0.3):
        whole = orc.advection(y, c, d, d, d)
        lo, hi = y[: n // 2], y[n // 2:]
        if c > 0:
            f_lo = orc.advection(lo, c, d, d, d, halo=hi[-1])
            f_hi = orc.advection(hi, c, d, d, d, halo=lo[-1])
            assert _same(f_lo, whole[: n // 2])
            assert _same(f_hi[1:], whole[n // 2 + 1:])
            np.testing.assert_allclose(f_hi[0], whole[n // 2], rtol=0, atol=1e-14)
        else:
            f_lo = orc.advection(lo, c, d, d, d, halo=hi[0])
            f_hi = orc.advection(hi, c, d, d, d, halo=lo[0])
            assert _same(f_hi, whole[n // 2:])
            assert _same(f_lo[:-1], whole[: n // 2 - 1])
            np.testing.assert_allclose(f_lo[-1], whole[n // 2 - 1], rtol=0, atol=1e-14)


def test_block_solve_inverts_the_newton_matrix():
    # property: (I - gamma J) x == b with J the analytic Jacobian of g
    rng = np.random.default_rng(5)
    n = 6
    p = orc.params()
    y = np.stack([1 + 0.2 * rng.uniform(-1, 1, (n, n, n)), 3.5 + 0.2 * rng.uniform(-1, 1, (n, n, n)),
                  3 + 0.2 * rng.uniform(-1, 1, (n, n, n))], axis=-1)
    b = rng.uniform(-1, 1, (n, n, n, 3))
    gamma = 1e-3
    x = orc.solve_reaction_linsys(y, b, gamma, p)
    u, v, w = y[..., 0], y[..., 1], y[..., 2]
    k2, k3, k4, k6 = p["k2"], p["k3"], p["k4"], p["k6"]
    J = np.zeros(y.shape[:-1] + (3, 3))
    J[..., 0, 0] = -k2 * w + 2 * k3 * u * v - k4
    J[..., 0, 1] = k3 * u * u
    J[..., 0, 2] = -k2 * u
    J[..., 1, 0] = k2 * w - 2 * k3 * u * v
    J[..., 1, 1] = -k3 * u * u
    J[..., 1, 2] = k2 * u
    J[..., 2, 0] = -k2 * w
    J[..., 2, 2] = -k2 * u - k6
    M = np.eye(3) - gamma * J
    r = np.einsum("...ij,...j->...i", M, x) - b
    assert np.abs(r).max() < 1e-10
