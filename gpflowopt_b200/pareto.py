"""Pareto front and cell decomposition of the non-dominated region: the host-side producer of the int32 cell
tables the HVPoI kernel consumes (mirror of gpflowopt/pareto.py; once per BO iteration, off the hot path).

``front`` / ``bounds.lb`` / ``bounds.ub`` are ndarray subclasses that also answer ``.value`` so code written
against the reference's DataHolders keeps working.
"""
import numpy as np

np_int_type = np.int32          # gpflow settings.dtypes.int_type
stability = 1e-6                # gpflow settings.numerics.jitter_level


class _Held(np.ndarray):
    @property
    def value(self):
        return np.asarray(self)


def _held(a, dtype=None):
    return np.asarray(a, dtype=dtype).view(_Held)


class BoundedVolumes(object):
    """Axis-aligned cells given by lower / upper index (or coordinate) rows."""

    @classmethod
    def empty(cls, dim, dtype):
        z = np.zeros((0, dim), dtype=dtype)
        return cls(z.copy(), z.copy())

    def __init__(self, lb, ub):
        lb, ub = np.atleast_2d(lb), np.atleast_2d(ub)
        assert lb.shape == ub.shape
        self.lb = _held(lb)
        self.ub = _held(ub)

    def append(self, lb, ub):
        self.lb = _held(np.vstack((self.lb, lb)), self.lb.dtype)
        self.ub = _held(np.vstack((self.ub, ub)), self.ub.dtype)

    def clear(self):
        dtype, outdim = self.lb.dtype, self.lb.shape[1]
        self.lb = _held(np.zeros((0, outdim), dtype=dtype))
        self.ub = _held(np.zeros((0, outdim), dtype=dtype))

    def size(self):
        return np.prod(np.asarray(self.ub) - np.asarray(self.lb), axis=1)


def non_dominated_sort(objectives, force_python=False):
    """(non-dominated rows, number of points dominating each row)  -- gpflowopt/pareto.py:78-90.
    Native host code (csrc/pareto_host.cu::gpfo_non_dominated_sort) unless ``force_python`` asks for the NumPy restatement
    below (kept reachable for the tests that pin both against the reference's known answers)."""
    Y = np.asarray(objectives)
    n = Y.shape[0]
    if not force_python and Y.ndim == 2 and n > 0 and 1 <= Y.shape[1] <= 64:
        from ._lib import non_dominated_counts
        dominance = non_dominated_counts(Y)
        return Y[dominance == 0], dominance
    dominance = np.zeros(n, dtype=np.int64)
    for k in range(n):                        # O(n^2 R) without the n x n x R temporary
        no_worse = np.all(Y[k] <= Y, axis=1)
        better = np.any(Y[k] < Y, axis=1)
        dominance += np.logical_and(no_worse, better)
    return Y[dominance == 0], dominance


class Pareto(object):
    def __init__(self, Y, threshold=0, force_python=False):
        self._force_python = force_python          # keep the NumPy reference implementation reachable for tests
        self.threshold = threshold
        self.Y = np.asarray(Y, dtype=np.float64)
        self.bounds = BoundedVolumes.empty(self.Y.shape[1], np_int_type)
        self.front = _held(np.zeros((0, self.Y.shape[1])))
        self.update()

    def _update_front(self):
        current = np.asarray(self.front)
        pf, _ = non_dominated_sort(self.Y)
        self.front = _held(pf[pf[:, 0].argsort(), :])         # ascending on the first objective
        return not np.array_equal(current, np.asarray(self.front))

    def update(self, Y=None, generic_strategy=False):
        self.Y = np.asarray(Y, dtype=np.float64) if Y is not None else self.Y
        if self._update_front():                               # cells only change with the front
            self.bounds.clear()
            if generic_strategy or self.Y.shape[1] != 2:
                self.divide_conquer_nd()
            else:
                self.bounds_2d()

    def bounds_2d(self):
        """Two objectives: P+1 strips, strip i spans ranks [i, i+1] in objective 0 and reaches from the ideal
        point up to the objective-1 value of the (P-i)-th point (gpflowopt/pareto.py:238-255)."""
        front = np.asarray(self.front)
        assert front.shape[1] == 2
        P = front.shape[0]
        rank_rows = np.vstack((np.zeros(2, dtype=np_int_type), np.argsort(front, axis=0) + 1,
                               np.full(2, P + 1, dtype=np_int_type)))
        i = np.arange(P + 1)
        lb = np.stack((i, np.zeros(P + 1, dtype=np.int64)), axis=1)
        ub = np.stack((i + 1, rank_rows[P + 1 - i, 1]), axis=1)
        self.bounds.lb = _held(lb, np_int_type)
        self.bounds.ub = _held(ub, np_int_type)

    def divide_conquer_nd(self):
        """Any number of objectives: binary space partitioning over the rank grid of the front; a cell is kept when
        its upper corner is not dominated, dropped when its lower corner is dominated, split otherwise
        (gpflowopt/pareto.py:173-236)."""
        front = np.asarray(self.front)
        P, R = front.shape
        if P > 0 and R <= 16 and not self._force_python:
            # native host code (csrc/pareto_host.cu): same visiting order, ~50x faster than the NumPy stack loop below
            from ._lib import pareto_cells
            lb, ub = pareto_cells(front, np.argsort(front, axis=0), self.threshold)
            if lb.shape[0]:
                self.bounds.lb = _held(lb, np_int_type)
                self.bounds.ub = _held(ub, np_int_type)
            return
        cols = np.arange(R)
        lo_pt = np.min(front, axis=0) - 1
        hi_pt = np.max(front, axis=0) + 1
        ext = np.vstack((lo_pt, front, hi_pt))
        rows = np.vstack((np.zeros(R, dtype=np_int_type), np.argsort(front, axis=0) + 1,
                          np.full(R, P + 1, dtype=np_int_type)))
        total = np.prod(hi_pt - lo_pt)

        def outside_dominated_region(corner):
            # the corner is strictly smaller than every front member in at least one objective
            return bool(np.all(np.any(corner < front, axis=1)))

        kept_lb, kept_ub = [], []
        todo = [(np.zeros(R, dtype=np_int_type), np.full(R, P + 1, dtype=np_int_type))]
        while todo:
            a, b = todo.pop()
            lb_rows, ub_rows = rows[a, cols], rows[b, cols]
            lb, ub = ext[lb_rows, cols], ext[ub_rows, cols]
            if outside_dominated_region(ub - stability):
                kept_lb.append(lb_rows)
                kept_ub.append(ub_rows)
            elif outside_dominated_region(lb + stability):
                width = b - a
                if np.any(width > 1) and np.all((np.prod(ub - lb) / total) > self.threshold):
                    k = int(np.argmax(width))
                    first = int(np.round(width[k] / 2.0))
                    lower_half_top = b.copy()
                    lower_half_top[k] -= first
                    upper_half_bottom = a.copy()
                    upper_half_bottom[k] += width[k] - first
                    todo.append((a.copy(), lower_half_top))
                    todo.append((upper_half_bottom, b.copy()))
        if kept_lb:
            self.bounds.lb = _held(np.vstack(kept_lb), np_int_type)
            self.bounds.ub = _held(np.vstack(kept_ub), np_int_type)

    def hypervolume(self, reference):
        """Hypervolume indicator of the dominated region (gpflowopt/pareto.py:257-282); reporting metric."""
        front = np.asarray(self.front)
        reference = np.asarray(reference, dtype=np.float64)
        min_pf = np.min(front, axis=0, keepdims=True)
        pseudo = np.vstack((min_pf, front, reference[None, :]))
        cols = np.arange(front.shape[1])
        ub = pseudo[np.asarray(self.bounds.ub), cols]
        lb = pseudo[np.asarray(self.bounds.lb), cols]
        return float(np.prod(reference[None, :] - min_pf) - np.sum(np.prod(ub - lb, axis=1)))


def merge_cells(pf_ext, lb, ub):
    """Coarsens a cell decomposition without changing the region it covers: cells that agree in all objectives but one
    and touch in that one (``a.ub[d] == b.lb[d]``, same row of ``pf_ext``) are fused, repeatedly, along every objective.

    The divide-and-conquer of the reference emits ~P^2 small cells (47 777 for a 204-point 3-objective front); fusing
    leaves ~P (1 736).  Both HVPoI terms are additive over such splits exactly -- Phi(c) - Phi(a) = (Phi(b) - Phi(a)) +
    (Phi(c) - Phi(b)) and (c - max(a, mu))+ = (b - max(a, mu))+ + (c - max(b, mu))+ -- so the acquisition value is the same
    function (differences are re-association roundings, ~1e-16 relative); the device kernel's cost is linear in the cell count.
    ``Pareto.bounds`` keeps the reference tables; only the copy uploaded to the device is coarsened.
    """
    lb = np.asarray(lb, dtype=np.int64)
    ub = np.asarray(ub, dtype=np.int64)
    if lb.shape[0] < 2:
        return lb.astype(np_int_type), ub.astype(np_int_type)
    R = lb.shape[1]
    pf_ext = np.asarray(pf_ext, dtype=np.float64)
    changed = True
    while changed:
        changed = False
        for d in range(R):
            others = [k for k in range(R) if k != d]
            keys = [pf_ext[lb[:, d], d]]                   # last key of lexsort is the primary one: others first, then d
            for k in reversed(others):
                keys += [ub[:, k], lb[:, k]]
            order = np.lexsort(keys)
            lb, ub = lb[order], ub[order]
            glue = np.all(lb[1:, others] == lb[:-1, others], axis=1) & np.all(ub[1:, others] == ub[:-1, others], axis=1) & \
                (lb[1:, d] == ub[:-1, d])
            if not glue.any():
                continue
            changed = True
            starts = np.flatnonzero(np.concatenate(([True], ~glue)))         # first cell of every run
            ends = np.concatenate((starts[1:] - 1, [lb.shape[0] - 1]))       # last cell of every run
            new_ub = ub[starts].copy()
            new_ub[:, d] = ub[ends, d]
            lb, ub = lb[starts], new_ub
    return lb.astype(np_int_type), ub.astype(np_int_type)
