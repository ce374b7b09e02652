"""Oracle: Pareto front and the cell decomposition of the non-dominated region (host-side setup of HVPoI).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  PINNED: reproduces the reference's known answers
(testing/unit/test_pareto.py:8-13, 43-57) and the golden tables in tests/golden/pareto_golden.npz, which
were produced by running the reference's own gpflowopt/pareto.py (tests/golden/gen_pareto_golden.py).

Restated reference sites (relative to /root/reference):
  * non_dominated_sort            gpflowopt/pareto.py:78-90
  * Pareto._update_front          gpflowopt/pareto.py:131-144  (front sorted ascending on objective 0)
  * Pareto.bounds_2d              gpflowopt/pareto.py:238-255
  * Pareto.divide_conquer_nd      gpflowopt/pareto.py:173-236  (stability = 1e-6, threshold = 0)
  * Pareto.hypervolume            gpflowopt/pareto.py:257-282  (indicator; only used to pin the tables)
"""
import numpy as np

STABILITY = 1e-6
INT = np.int32                      # gpflow settings.dtypes.int_type (pareto.py:21)


def non_dominated_sort(objectives):
    """Returns (non-dominated rows, number of dominating points per row)."""
    Y = np.asarray(objectives)
    # row i is dominated by row k iff Y[k] <= Y[i] everywhere and Y[k] < Y[i] somewhere
    le = np.all(Y[None, :, :] <= Y[:, None, :], axis=2)
    lt = np.any(Y[None, :, :] < Y[:, None, :], axis=2)
    dominance = np.sum(np.logical_and(le, lt), axis=1)
    return Y[dominance == 0], dominance


def sorted_front(Y):
    pf, _ = non_dominated_sort(Y)
    return pf[pf[:, 0].argsort(), :]


def _augments_or_dominates(smaller):
    """True iff the test point is smaller in at least one objective than every front member."""
    return bool(np.all(np.any(smaller, axis=1)))


def bounds_2d(front):
    P = front.shape[0]
    order = np.argsort(front, axis=0)
    ext_idx = np.vstack((np.zeros(2, dtype=INT), order + 1, np.ones(2, dtype=INT) * P + 1))
    lb = np.zeros((P + 1, 2), dtype=INT)
    ub = np.zeros((P + 1, 2), dtype=INT)
    for i in range(P + 1):
        lb[i] = (i, 0)
        ub[i] = (i + 1, ext_idx[-i - 1, 1])
    return lb, ub


def divide_conquer_nd(front, threshold=0.0):
    P, R = front.shape
    dims = np.arange(R)
    order = np.argsort(front, axis=0) + 1                       # 0 is reserved for the ideal point
    lo = np.min(front, axis=0) - 1
    hi = np.max(front, axis=0) + 1
    ext = np.vstack((lo, front, hi))
    ext_idx = np.vstack((np.zeros(R, dtype=INT), order, np.ones(R, dtype=INT) * P + 1))
    total = np.prod(hi - lo)
    lbs, ubs = [], []
    stack = [(np.zeros(R, dtype=INT), (ext_idx.shape[0] - 1) * np.ones(R, dtype=INT))]
    while stack:
        c_lo, c_hi = stack.pop()
        lb = ext[ext_idx[c_lo, dims], dims]
        ub = ext[ext_idx[c_hi, dims], dims]
        if _augments_or_dominates((ub - STABILITY) < front):
            lbs.append(ext_idx[c_lo, dims])
            ubs.append(ext_idx[c_hi, dims])
        elif _augments_or_dominates((lb + STABILITY) < front):
            span = c_hi - c_lo
            if np.any(span > 1) and np.all((np.prod(ub - lb) / total) > threshold):
                k = int(np.argmax(span))
                edge = span[k]
                first = int(np.round(edge / 2.0))
                second = edge - first
                upper = np.copy(c_hi)
                upper[k] -= first
                stack.append((np.copy(c_lo), upper))
                lower = np.copy(c_lo)
                lower[k] += second
                stack.append((lower, np.copy(c_hi)))
    if not lbs:
        return np.zeros((0, R), dtype=INT), np.zeros((0, R), dtype=INT)
    return np.vstack(lbs).astype(INT), np.vstack(ubs).astype(INT)


class Pareto:
    def __init__(self, Y, threshold=0.0):
        self.threshold = threshold
        self.Y = np.asarray(Y, dtype=np.float64)
        R = self.Y.shape[1]
        self.front = np.zeros((0, R))
        self.lb = np.zeros((0, R), dtype=INT)
        self.ub = np.zeros((0, R), dtype=INT)
        self.update()

    def update(self, Y=None, generic_strategy=False):
        if Y is not None:
            self.Y = np.asarray(Y, dtype=np.float64)
        new_front = sorted_front(self.Y)
        changed = not np.array_equal(self.front, new_front)
        self.front = new_front
        if changed:                                             # pareto.py:161-171
            if generic_strategy or self.Y.shape[1] != 2:
                self.lb, self.ub = divide_conquer_nd(self.front, self.threshold)
            else:
                self.lb, self.ub = bounds_2d(self.front)
        return changed

    def hypervolume(self, reference):
        reference = np.asarray(reference, dtype=np.float64)
        min_pf = np.min(self.front, axis=0, keepdims=True)
        pseudo = np.vstack((min_pf, self.front, reference[None, :]))
        col = np.arange(pseudo.shape[1])
        ub = pseudo[self.ub, col]
        lb = pseudo[self.lb, col]
        return float(np.prod(reference[None, :] - min_pf) - np.sum(np.prod(ub - lb, axis=1)))
