"""Expected Improvement (mirror of gpflowopt/acquisition/ei.py)."""
import numpy as np

from .acquisition import Acquisition


class ExpectedImprovement(Acquisition):
    def __init__(self, model):
        super(ExpectedImprovement, self).__init__(model)
        self.fmin = np.zeros(1)
        self._setup()

    def _setup(self):
        super(ExpectedImprovement, self)._setup()
        # lowest posterior mean over the feasible training inputs (gpflowopt/acquisition/ei.py:63-68)
        feasible_samples = self.data[0][self.highest_parent.feasible_data_index(), :]
        samples_mean, _ = self.models[0].predict_f(feasible_samples)
        self.fmin = np.min(samples_mean, axis=0)

    def build_acquisition(self, Xcand):
        return Xcand.ei(self.models[0], self.fmin)
