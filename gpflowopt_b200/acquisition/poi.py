"""Probability of Improvement (mirror of gpflowopt/acquisition/poi.py)."""
import numpy as np

from .acquisition import Acquisition


class ProbabilityOfImprovement(Acquisition):
    def __init__(self, model):
        super(ProbabilityOfImprovement, self).__init__(model)
        self.fmin = np.zeros(1)
        self._setup()

    def _setup(self):
        super(ProbabilityOfImprovement, self)._setup()
        feasible_samples = self.data[0][self.highest_parent.feasible_data_index(), :]     # poi.py:57-61
        samples_mean, _ = self.models[0].predict_f(feasible_samples)
        self.fmin = np.min(samples_mean, axis=0)

    def build_acquisition(self, Xcand):
        return Xcand.poi(self.models[0], self.fmin)
