"""CPU: the LSAP restatement (oracle/lsap_oracle.py) against scipy.optimize.linear_sum_assignment, the solver the
reference calls (matcher.py:151) -- identical INDICES, including on tie-heavy integer matrices where several optima
exist, with both the sequential selection rule and the order-independent key the device kernel reduces with."""
import numpy as np
import pytest
from scipy.optimize import linear_sum_assignment

from oracle import lsap_oracle as L


def _cases():
    rng = np.random.default_rng(11)
    out = []
    for (r, c) in ((1, 1), (1, 7), (7, 1), (5, 5), (20, 6), (6, 20), (100, 13), (13, 100), (40, 40), (100, 37)):
        out.append(("float", rng.standard_normal((r, c)).astype(np.float32)))
        out.append(("ties", rng.integers(0, 3, (r, c)).astype(np.float32)))
        out.append(("ties2", rng.integers(0, 2, (r, c)).astype(np.float64)))
    out.append(("const", np.ones((6, 6), np.float32)))
    out.append(("const_wide", np.zeros((3, 9), np.float32)))
    out.append(("const_tall", np.zeros((9, 3), np.float32)))
    out.append(("inf_ok", np.array([[np.inf, 1.0], [2.0, np.inf]])))
    return out


@pytest.mark.parametrize("parallel_rule", [False, True])
def test_lsap_oracle_equals_scipy(parallel_rule):
    for name, C in _cases():
        ri, ci = linear_sum_assignment(C)
        oi, oj = L.solve(C, parallel_rule=parallel_rule)
        assert np.array_equal(ri, oi) and np.array_equal(ci, oj), (name, C.shape)


def test_lsap_oracle_empty_and_invalid():
    for shape in ((0, 4), (4, 0), (0, 0)):
        i, j = L.solve(np.zeros(shape))
        assert len(i) == 0 and len(j) == 0
    with pytest.raises(ValueError):
        L.solve(np.array([[np.nan, 1.0]]))
    with pytest.raises(ValueError):
        L.solve(np.array([[np.inf, np.inf], [1.0, 2.0]]))


def test_lsap_oracle_fuzz_small_integer_matrices():
    """400 random small matrices with few distinct values (many optima): both selection rules return scipy's pairs."""
    rng = np.random.default_rng(123)
    for _ in range(400):
        r, c = int(rng.integers(1, 9)), int(rng.integers(1, 9))
        C = rng.integers(0, int(rng.integers(1, 4)) + 1, (r, c)).astype(np.float64)
        ri, ci = linear_sum_assignment(C)
        for rule in (False, True):
            oi, oj = L.solve(C, parallel_rule=rule)
            assert np.array_equal(ri, oi) and np.array_equal(ci, oj), (C, rule)
