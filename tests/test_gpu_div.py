"""The row solve divides by diagonal entries of J M^-1 J^T (SolverBoxPGS.cpp:100-112) with the fast path of div.rn.f32
written out (pgs_math.cuh: div_rn_nocall), so that no subroutine CALL sits in the sweep loop. This pins the claim the
kernels rely on: for normal operands with a normal quotient the result is the correctly rounded quotient, bit for bit
what the reference's `x / A` gives on the host; zero and infinite numerators keep their IEEE result."""
import numpy as np
import pytest

from slrbs_b200 import capi

pytestmark = pytest.mark.gpu


def bits(x):
    return np.asarray(x, np.float32).view(np.uint32)


def test_solver_range_is_ieee_division_bit_for_bit():
    rng = np.random.default_rng(11)
    n = 1 << 20
    # numerators: impulses / right-hand sides over twelve decades, both signs; denominators: positive diagonal entries
    a = (rng.standard_normal(n) * 10.0 ** rng.uniform(-8, 4, n)).astype(np.float32)
    b = (10.0 ** rng.uniform(-4, 4, n)).astype(np.float32)
    got = capi.debug_div_rn(a, b)
    want = (a / b).astype(np.float32)          # numpy float32 division is IEEE round-to-nearest
    assert np.array_equal(bits(got), bits(want)), int(np.count_nonzero(bits(got) != bits(want)))


def test_time_step_denominator():
    rng = np.random.default_rng(12)
    a = (rng.standard_normal(1 << 18) * 10.0 ** rng.uniform(-6, 6, 1 << 18)).astype(np.float32)
    for dt in (0.01, 1.0 / 60.0, 0.002):
        b = np.full_like(a, np.float32(dt))
        assert np.array_equal(bits(capi.debug_div_rn(a, b)), bits((a / b).astype(np.float32)))


def test_zero_infinite_and_nan_numerators():
    a = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1.0], np.float32)
    b = np.array([2.0, 3.0, 0.5, 0.25, 1.0, 4.0], np.float32)
    got = capi.debug_div_rn(a, b)
    with np.errstate(all="ignore"):
        want = (a / b).astype(np.float32)
    assert np.array_equal(bits(got[:4]), bits(want[:4])) and np.isnan(got[4]) and got[5] == np.float32(0.25)
