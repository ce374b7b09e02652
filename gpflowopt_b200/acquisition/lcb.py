"""Lower Confidence Bound (mirror of gpflowopt/acquisition/lcb.py)."""
import numpy as np

from .acquisition import Acquisition


class LowerConfidenceBound(Acquisition):
    def __init__(self, model, sigma=2.0):
        super(LowerConfidenceBound, self).__init__(model)
        self.sigma = np.array(sigma, dtype=np.float64)

    def build_acquisition(self, Xcand):
        return Xcand.lcb(self.models[0], self.sigma)
