"""oracle/cvfused_oracle.py -- TEST INFRASTRUCTURE, not product code.

numpy restatement of the element-wise arithmetic of CVODE's seven fused-kernel plugin functions, in the
operation order of the reference's CPU implementation (src/cvode/cvode_fused_stubs.c, line ranges at each
function) INCLUDING the branch of N_VLinearSum's case analysis (src/nvector/serial/nvector_serial.c:397-477)
the scalars select.  Every numpy operation rounds once (no FMA), like the reference's C compiled without
contraction.  Pinned bit for bit against the reference stubs running on nvector_serial by
tests/test_cvode_fused_oracle_cpu.py; the CUDA functors of sundials_b200/csrc/b200vec_cvfused.cu are
written from the same formulas and checked against the stubs directly on the GPU
(tests/test_cvode_fused_gpu.py).
"""
import numpy as np

FRACT = 0.1  # cvode_fused_stubs.c:28


def ewt(rtol, atol, y):
    """stubs:38-72 -> (tempv, weight); atol scalar or array"""
    t = (rtol * np.abs(y)) + atol
    return t, 1.0 / t


def constraints(c, ewt_, y, mm):
    """stubs:80-89"""
    t = np.where(np.abs(c) >= 1.5, 1.0, 0.0)
    t = t * c
    t = t / ewt_
    t = (-0.1 * t) + y          # N_VLinearSum(1, y, -0.1, tmp, tmp): VLin1
    return t * mm


def nls_resid(rl1, ngamma, zn1, ycor, ftemp):
    """stubs:97-104"""
    t = (rl1 * zn1) + ycor
    return (ngamma * ftemp) + t


def diag_form_y(h, r, fpred, zn1, ypred):
    """stubs:112-119 -> (ftemp, y)"""
    f = (h * fpred) - zn1
    return f, (r * f) + ypred


def diag_build_m(uround, h, ftemp, fpred, ewt_, M):
    """stubs:128-147 -> (bit, bitcomp, y, M)"""
    a, b = FRACT, -h
    M = M - fpred
    if b in (1.0, -1.0):
        M = (a * ftemp) + (b * M)
    elif a == b:
        M = a * (ftemp + M)      # VScaleSum
    elif a == -b:
        M = a * (ftemp - M)      # VScaleDiff
    else:
        M = (a * ftemp) + (b * M)
    y = ftemp * ewt_
    bit = np.where(np.abs(y) >= uround, 1.0, 0.0)
    bc = bit + (-1.0)
    y = ftemp * bit
    y = (FRACT * y) - bc
    M = M / y
    M = M * bit
    M = M - bc
    return bit, bc, y, M


def diag_update_m(r, M):
    """stubs:154-161"""
    m = 1.0 / M
    m = m + (-1.0)
    m = r * m
    return m + 1.0
