"""ar3d_oracle.py -- TEST INFRASTRUCTURE: CPU restatement (numpy) of the right-hand-side
functions of the reference's benchmarks/advection_reaction_3D/raja, the checker of
apps/advection_reaction_3D.  Only tests/, __graft_entry__.smoke() and bench.py's CPU leg may
import it; the product never does.

Every function states the reference lines it follows.  numpy element-wise arithmetic is IEEE
double without contraction, and every expression below keeps the reference's operand order, so
the results are bit-identical to the reference's CPU build: PINNED by
tests/test_ar3d_oracle_cpu.py against tests/golden/advection_reaction_3D/rhs_*.npz, which hold
the outputs of the reference's own functions run here (tests/c/ar3d_rhs_dump.cpp,
tests/golden/make_ar3d_golden.py).

State layout: y[i, j, k, l] (x slowest, species l = u, v, w fastest), RAJA::Layout<4>(nxl, nyl,
nzl, dof) of rhs3D.hpp:62.  One rank (periodic in all directions) unless `halo` is given: then
`halo` is the plane of the upstream x-neighbour (ny, nz, 3) -- the contents of the
reference's Wrecv (c > 0) or Erecv (c < 0) buffer.
"""
from __future__ import annotations

import math

import numpy as np

DEFAULTS = dict(xmax=1.0, A=1.0, B=3.5, k1=1.0, k2=1.0, k3=1.0, k4=1.0, k5=1.0 / 5.0e-6, k6=1.0 / 5.0e-6, c=0.01)


def params(**kw):
    p = dict(DEFAULTS)
    p.update(kw)
    return p


def advection(y, c, dx, dy, dz, halo=None):
    """Advection, rhs3D.hpp:30-319: first-order upwind differences, periodic.

    Interior points (local i, j, k >= 1 for c > 0) sum the z, y, x terms in that order
    (:82-87); the points of the three upstream faces sum x, y, z (:182-187, :204-209, :226-231).
    c == 0 leaves the zero that N_VConst wrote (:59)."""
    y = np.asarray(y, dtype=np.float64)
    nx, ny, nz, _ = y.shape
    out = np.zeros_like(y)
    if c == 0.0:
        return out
    cx, cy, cz = -c / dx, -c / dy, -c / dz
    sh = 1 if c > 0.0 else -1
    yi = np.roll(y, sh, axis=0)
    yj = np.roll(y, sh, axis=1)
    yk = np.roll(y, sh, axis=2)
    if halo is not None:
        yi[0 if c > 0.0 else nx - 1] = halo
    if c > 0.0:
        tz, ty, tx = cz * (y - yk), cy * (y - yj), cx * (y - yi)
    else:
        tz, ty, tx = cz * (yk - y), cy * (yj - y), cx * (yi - y)
    interior = (tz + ty) + tx
    face = (tx + ty) + tz
    m = np.zeros((nx, ny, nz, 1), dtype=bool)
    e = 0 if c > 0.0 else -1
    m[e, :, :] = True
    m[:, e, :] = True
    m[:, :, e] = True
    return np.where(m, face, interior)


def reaction_terms(y, p):
    """g(y), rhs3D.hpp:373-378 (operand order of the reference's expressions)."""
    u, v, w = y[..., 0], y[..., 1], y[..., 2]
    A, B = p["A"], p["B"]
    k1, k2, k3, k4, k5, k6 = (p[k] for k in ("k1", "k2", "k3", "k4", "k5", "k6"))
    g = np.empty_like(y)
    g[..., 0] = k1 * A - k2 * w * u + k3 * u * u * v - k4 * u
    g[..., 1] = k2 * w * u - k3 * u * u * v
    g[..., 2] = -k2 * w * u + k5 * B - k6 * w
    return g


def reaction(y, p, into=None):
    """Reaction, rhs3D.hpp:322-383: ydot (zeroed unless add_reactions) += g(y)."""
    base = np.zeros_like(y) if into is None else into
    return base + reaction_terms(np.asarray(y, dtype=np.float64), p)


def advection_reaction(y, p, dx, dy, dz, halo=None):
    """AdvectionReaction, rhs3D.hpp:386-406: advection first, reactions added to it."""
    return reaction(y, p, into=advection(y, p["c"], dx, dy, dz, halo))


def solve_reaction_linsys(y, b, gamma, p):
    """SolveReactionLinSys, rhs3D.hpp:441-550: x = (I - gamma dg/dy)^-1 b per node, the closed
    form with the reference's intermediate products."""
    u, v, w = y[..., 0], y[..., 1], y[..., 2]
    b0, b1, b2 = b[..., 0], b[..., 1], b[..., 2]
    k2, k3, k4, k6 = p["k2"], p["k3"], p["k4"], p["k6"]
    A0 = -k2 * w + 2.0 * k3 * u * v - k4
    A1 = k3 * u * u
    A2 = -k2 * u
    A3 = k2 * w - 2.0 * k3 * u * v
    A4 = -k3 * u * u
    A5 = k2 * u
    A6 = -k2 * w
    A7 = np.zeros_like(u)
    A8 = -k2 * u - k6
    A0 = 1.0 - (gamma * A0)
    A1 = -gamma * A1
    A2 = -gamma * A2
    A3 = -gamma * A3
    A4 = 1.0 - (gamma * A4)
    A5 = -gamma * A5
    A6 = -gamma * A6
    A7 = -gamma * A7
    A8 = 1.0 - (gamma * A8)
    s0 = A4 * A8
    s1 = A1 * A5
    s2 = A2 * A7
    s3 = A5 * A7
    s4 = A1 * A8
    s5 = A2 * A4
    s6 = 1.0 / (A0 * s0 - A0 * s3 + A3 * s2 - A3 * s4 + A6 * s1 - A6 * s5)
    s7 = A2 * A3
    s8 = A6 * b0
    s9 = A2 * A6
    s10 = A3 * b0
    s11 = 1.0 / A0
    s12 = A1 * s11
    s13 = (-A6 * s12 + A7) / (-A3 * s12 + A4)
    x = np.empty_like(y)
    x[..., 0] = s6 * (b0 * (s0 - s3) + b1 * (s2 - s4) + b2 * (s1 - s5))
    x[..., 1] = s6 * (b2 * (s7 - A0 * A5) + b1 * (A0 * A8 - s9) + A5 * s8 - A8 * s10)
    x[..., 2] = (-b2 + s11 * s8 + s13 * (b1 - s10 * s11)) / (-A8 + s11 * s9 + s13 * (A5 - s11 * s7))
    return x


def gaussian_factor(coord, xmax):
    """one axis of Gaussian3D, advection_reaction_3D.cpp:539-557 (libm exp / pow / sqrt)."""
    alpha = 0.1
    mu = xmax / 2.0
    sigma = xmax / 4.0
    denom = 2.0 * math.sqrt((sigma * sigma * sigma) * math.pow(2 * math.pi, 3))
    return alpha * math.exp(-((coord - mu) * (coord - mu) * (1.0 / sigma)) / denom)


def initial_condition(npts, p, rank=0, nranks=1):
    """SetIC, advection_reaction_3D.cpp:560-616, for the slab of `rank` (dims = {nranks, 1, 1})."""
    n = int(npts)
    i0 = n * rank // nranks
    nxl = n * (rank + 1) // nranks - i0
    d = p["xmax"] / float(n)
    gx = np.array([gaussian_factor((rank * nxl + i) * d, p["xmax"]) for i in range(nxl)])
    gy = np.array([gaussian_factor(j * d, p["xmax"]) for j in range(n)])
    gz = gy.copy()
    pert = (gx[:, None, None] + gy[None, :, None]) + gz[None, None, :]
    us = p["k1"] * p["A"] / p["k4"]
    vs = p["k2"] * p["k4"] * p["B"] / (p["k1"] * p["k3"] * p["A"])
    ws = 3.0
    y = np.empty((nxl, n, n, 3))
    y[..., 0] = us + pert
    y[..., 1] = vs + pert
    y[..., 2] = ws + pert
    return y
