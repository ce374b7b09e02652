"""Hypervolume-based Probability of Improvement (mirror of gpflowopt/acquisition/hvpoi.py)."""
import numpy as np

from ..pareto import Pareto, merge_cells
from .acquisition import Acquisition


class HVProbabilityOfImprovement(Acquisition):
    def __init__(self, models):
        super(HVProbabilityOfImprovement, self).__init__(models)
        num_objectives = self.data[1].shape[1]
        assert num_objectives > 1
        self.pareto = Pareto(np.empty((0, num_objectives)))
        self.reference = np.ones((1, num_objectives))
        self._cells_version = 0
        self._merged, self._merged_version = None, -1

    def _estimate_reference(self):
        pf = self.pareto.front
        f = np.max(pf, axis=0, keepdims=True) - np.min(pf, axis=0, keepdims=True)
        return np.max(pf, axis=0, keepdims=True) + 2 * f / pf.shape[0]

    def _setup(self):
        """Pareto set of the predicted means and the cells covering the non-dominated region
        (gpflowopt/acquisition/hvpoi.py:84-96)."""
        super(HVProbabilityOfImprovement, self)._setup()
        feasible_samples = self.data[0][self.highest_parent.feasible_data_index(), :]
        F = np.hstack([m.predict_f(feasible_samples)[0] for m in self.models])
        self.pareto.update(F)
        self.reference = self._estimate_reference()
        self._cells_version += 1

    def build_acquisition(self, Xcand):
        outdim = self.data[1].shape[1]
        pf_ext = np.vstack((-np.inf * np.ones((1, outdim)), self.pareto.front, self.reference))
        if self._merged_version != self._cells_version:     # coarsen once per Pareto update (same region, fewer cells)
            self._merged = merge_cells(pf_ext, self.pareto.bounds.lb, self.pareto.bounds.ub)
            self._merged_version = self._cells_version
        return Xcand.hvpoi(self.models, pf_ext, self._merged[0], self._merged[1], (id(self), self._cells_version))
