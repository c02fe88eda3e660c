"""CPU suite (no GPU): pins the oracle (oracle/nvec_oracle.c) against

  1. the UNMODIFIED reference nvector_serial (oracle/_ref, when present),
     bit-for-bit on every parity case at several lengths;
  2. the committed golden vectors generated from that reference
     (tests/golden/nvec_golden.npz, tests/golden/make_golden.py);
  3. the known answers of the reference's own unit tests
     (test/unit_tests/nvector/test_nvector.c).
"""
from pathlib import Path

import numpy as np
import pytest

from _cases import all_cases, compare

CASES = all_cases()
GOLDEN = Path(__file__).resolve().parent / "golden" / "nvec_golden.npz"
GOLDEN_SEED = 20240607


@pytest.mark.parametrize("n", [1, 7, 200, 1027])
def test_oracle_matches_reference_serial_bitwise(oracle, refserial, n):
    for name, fn in CASES:
        got = fn(oracle, n, 1000 + n)
        want = fn(refserial, n, 1000 + n)
        compare(f"{name}@n={n}", got, want, exact_reductions=True, n=n)


def test_oracle_matches_golden_vectors(oracle):
    g = np.load(GOLDEN)
    keys = set(g.files)
    checked = 0
    for n in (67, 1):
        for name, fn in CASES:
            res = fn(oracle, n, GOLDEN_SEED)
            for k, v in res.items():
                key = f"n{n}|{name}|{k}"
                if key not in keys:
                    continue
                got = np.atleast_1d(np.asarray(v, dtype=np.float64))
                want = g[key]
                assert got.shape == want.shape, key
                assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), key
                checked += 1
    assert checked == len(keys) and checked > 800


# ---- known answers from the reference's unit tests (exactly representable data)
def test_known_answers_test_nvector_c(oracle):
    n = 1000
    two, half, one, neg1 = np.full(n, 2.0), np.full(n, 0.5), np.full(n, 1.0), np.full(n, -1.0)
    # Test_N_VDotProd (test_nvector.c:1395-1415): x=2, y=1/2 -> global length
    assert oracle.dot_prod(two, half) == n
    # Test_N_VMaxNorm (:1428-1460): x=-1/2 except one -1 -> 1
    x = np.full(n, -0.5); x[n - 1] = -1.0
    assert oracle.max_norm(x) == 1.0
    # Test_N_VWrmsNorm (:1470-1500): x=-1/2, w=1/2 -> 1/4
    assert oracle.wrms_norm(np.full(n, -0.5), half) == 0.25
    # Test_N_VWrmsNormMask (:1510-1550): id = 1 except last 0 -> 1/4*sqrt((n-1)/n)
    id = np.ones(n); id[n - 1] = 0.0
    assert abs(oracle.wrms_norm_mask(np.full(n, -0.5), half, id) - 0.25 * np.sqrt((n - 1) / n)) < 1e-15
    # Test_N_VMin (:1560-1590)
    x = np.full(n, 2.0); x[n - 1] = -2.0
    assert oracle.min(x) == -2.0
    # Test_N_VWL2Norm (:1600-1630): 1/4*sqrt(n)
    assert abs(oracle.wl2_norm(np.full(n, -0.5), half) - 0.25 * np.sqrt(n)) < 1e-13
    # Test_N_VL1Norm (:1640-1670): x=-1 -> n
    assert oracle.l1_norm(neg1) == n
    # Test_N_VInvTest (:1862-1897): zeros leave z untouched and return false
    x = np.where(np.arange(n) % 2 == 0, 0.0, 0.5); z = np.zeros(n)
    assert oracle.inv_test(x, z) is False
    assert np.all(z[0::2] == 0.0) and np.all(z[1::2] == 2.0)
    z = np.zeros(n)
    assert oracle.inv_test(half, z) is True and np.all(z == 2.0)
    # Test_N_VConstrMask (:1910-1990): 7 cases cycling
    c = np.zeros(n); xx = np.zeros(n)
    pat = [(-2.0, -2.0), (-1.0, -1.0), (-1.0, 0.0), (0.0, 0.5), (1.0, 0.0), (1.0, 1.0), (2.0, 2.0)]
    for i in range(n):
        c[i], xx[i] = pat[i % 7]
    m = np.full(n, 5.0)
    assert oracle.constr_mask(c, xx, m) is True and np.all(m == 0.0)
    pat = [(-2.0, 2.0), (-2.0, 0.0), (-1.0, 2.0), (1.0, -2.0), (2.0, 0.0), (2.0, -2.0), (0.0, -1.0)]
    for i in range(n):
        c[i], xx[i] = pat[i % 7]
    assert oracle.constr_mask(c, xx, m) is False
    assert np.array_equal(m, np.where(np.arange(n) % 7 == 6, 0.0, 1.0))
    # Test_N_VMinQuotient (:2100-2135): num=2, denom=2 -> 1; denom=0 -> SUN_BIG_REAL
    assert oracle.min_quotient(two, two) == 1.0
    assert oracle.min_quotient(two, np.zeros(n)) == np.finfo(np.float64).max
    # Test_N_VLinearSum case 1a (:575-600): y = x + y in place, x=1, y=-2 -> -1
    y = np.full(n, -2.0)
    oracle.linear_sum(1.0, one, 1.0, y, y)
    assert np.all(y == -1.0)


def test_linear_sum_form_selection(oracle):
    # serial:397-465 decision order
    F = oracle.linear_sum_form
    assert F(3.0, 1.0, False, True) == 0      # axpy into y
    assert F(1.0, 3.0, True, False) == 1      # axpy into x
    assert F(1.0, 1.0, False, False) == 2     # sum
    assert F(-1.0, 1.0, False, False) == 3    # y - x
    assert F(1.0, -1.0, False, False) == 4    # x - y
    assert F(1.0, 0.0, False, False) == 5     # lin1 (b*y + x)
    assert F(0.5, 1.0, False, False) == 6
    assert F(-1.0, 0.5, False, False) == 7
    assert F(0.5, -1.0, False, False) == 8
    assert F(0.5, 0.5, False, False) == 9
    assert F(0.5, -0.5, False, False) == 10
    assert F(0.5, 0.25, False, False) == 11
    assert F(1.0, 1.0, True, True) == 0       # first test wins


def test_mpi_wrms_semantics(oracle):
    # nvector_manyvector.c:940-965: sqrt(sum of local sums / GLOBAL length)
    rng = np.random.default_rng(5)
    x, w = rng.uniform(-1, 1, 1000), rng.uniform(0.5, 2, 1000)
    parts = [oracle.wsqr_sum(x[a:b], w[a:b]) for a, b in ((0, 300), (300, 1000))]
    got = oracle.mpi_wrms_from_local(parts, 1000)
    assert abs(got - oracle.wrms_norm(x, w)) <= 1e-15 * got


def test_fill_uniform_lcg(oracle):
    x = oracle.fill_uniform(5, 12345)
    s, want = 12345, []
    for _ in range(5):
        s = (1103515245 * s + 12345) & 0x7FFFFFFF
        want.append(2.0 * (s / 0x7FFFFFFF) - 1.0)
    assert np.array_equal(x, np.array(want))
