"""Probability of Feasibility (mirror of gpflowopt/acquisition/pof.py)."""
import numpy as np

from .acquisition import Acquisition


class ProbabilityOfFeasibility(Acquisition):
    def __init__(self, model, threshold=0.0, minimum_pof=0.5):
        super(ProbabilityOfFeasibility, self).__init__(model)
        self.threshold = threshold
        self.minimum_pof = minimum_pof

    def constraint_indices(self):
        return np.arange(self.data[1].shape[1])

    def feasible_data_index(self):
        """Model-based feasibility of the training points (gpflowopt/acquisition/pof.py:63-79)."""
        pred = self.evaluate(self.data[0])
        return pred.ravel() > self.minimum_pof

    def build_acquisition(self, Xcand):
        return Xcand.pof(self.models[0], self.threshold)
