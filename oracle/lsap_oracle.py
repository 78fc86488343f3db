"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the rectangular linear-sum-assignment solver the reference calls
(``scipy.optimize.linear_sum_assignment``, reference mask2former/modeling/matcher.py:8,151; scipy is a third-party
dependency of the reference, unpinned in its requirements.txt; the container has scipy 1.18, which is the pin here).

The algorithm is the shortest-augmenting-path method of D. F. Crouse, "On implementing 2D rectangular assignment
algorithms", IEEE T-AES 52(4), 2016, as published in scipy (``scipy/optimize/rectangular_lsap``): rows are added one
at a time, each by a Dijkstra-like scan over the columns not yet in the tree; dual variables ``u, v`` keep reduced
costs non-negative; a wide matrix is solved as is, a tall one through its transpose.  Two details decide WHICH optimum
is returned when several exist, and are restated here because the device kernel (``mp_former_b200/csrc/matcher.cu``)
must reproduce them to return identical indices:
  * the scan order: ``remaining`` starts as ``nc-1, nc-2, ..., 0`` and a visited column is replaced by the last one;
  * among columns of equal tentative distance an UNASSIGNED column wins (the last such in scan order), otherwise the
    first in scan order.
``scan_key`` is the order-independent form of that rule used by the kernel's parallel reduction; ``solve(...,
parallel_rule=True)`` runs the restatement with it so the CPU tests pin the equivalence.

Pinned by ``tests/test_lsap_oracle_cpu.py`` against scipy itself (random float, integer/tie-heavy, constant, tall,
wide, empty and single-row/column matrices).
"""
import math

import numpy as np


def scan_key(value, it, unassigned):
    """Total order whose minimum is the column scipy's sequential scan selects."""
    return (value, 0, -it) if unassigned else (value, 1, it)


def solve(cost, parallel_rule=False):
    """cost: 2-D array-like [rows, cols] -> (row_ind, col_ind) int64 arrays, like scipy."""
    C = np.asarray(cost, dtype=np.float64)
    if C.ndim != 2:
        raise ValueError("expected a matrix")
    R, Cn = C.shape
    if R == 0 or Cn == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    transposed = Cn < R
    if transposed:
        C = C.T.copy()
    nr, nc = C.shape
    if np.isnan(C).any() or np.isneginf(C).any():
        raise ValueError("matrix contains invalid numeric entries")
    u = [0.0] * nr
    v = [0.0] * nc
    path = [-1] * nc
    col4row = [-1] * nr
    row4col = [-1] * nc
    for cur in range(nr):
        spc = [math.inf] * nc
        SR = [False] * nr
        SC = [False] * nc
        remaining = [nc - it - 1 for it in range(nc)]
        num_remaining = nc
        min_val = 0.0
        i = cur
        sink = -1
        while sink == -1:
            SR[i] = True
            index = -1
            lowest = math.inf
            best = None
            for it in range(num_remaining):
                j = remaining[it]
                r = min_val + C[i, j] - u[i] - v[j]
                if r < spc[j]:
                    path[j] = i
                    spc[j] = r
                if parallel_rule:
                    if spc[j] < math.inf:
                        k = scan_key(spc[j], it, row4col[j] == -1)
                        if best is None or k < best:
                            best = k
                            index, lowest = it, spc[j]
                elif spc[j] < lowest or (spc[j] == lowest and row4col[j] == -1):
                    lowest = spc[j]
                    index = it
            min_val = lowest
            if min_val == math.inf:
                raise ValueError("cost matrix is infeasible")
            j = remaining[index]
            if row4col[j] == -1:
                sink = j
            else:
                i = row4col[j]
            SC[j] = True
            num_remaining -= 1
            remaining[index] = remaining[num_remaining]
        u[cur] += min_val
        for r_ in range(nr):
            if SR[r_] and r_ != cur:
                u[r_] += min_val - spc[col4row[r_]]
        for j in range(nc):
            if SC[j]:
                v[j] -= min_val - spc[j]
        j = sink
        while True:
            i = path[j]
            row4col[j] = i
            col4row[i], j = j, col4row[i]
            if i == cur:
                break
    if transposed:
        order = np.argsort(np.asarray(col4row, dtype=np.int64), kind="stable")
        return np.asarray(col4row, dtype=np.int64)[order], order.astype(np.int64)
    return np.arange(nr, dtype=np.int64), np.asarray(col4row, dtype=np.int64)
