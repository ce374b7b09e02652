"""Min-value entropy search (mirror of gpflowopt/acquisition/mes.py; Wang & Jegelka 2017).

Host side = the Gumbel approximation of the distribution of the minimum y* (reference mes.py:63-94); device side = one
GPFO_OP_MES leaf holding the sampled minima (reference mes.py:96-102)."""
import numpy as np
from scipy.optimize import bisect
from scipy.stats import norm

from ..design import RandomDesign
from .acquisition import Acquisition


class MinValueEntropySearch(Acquisition):
    def __init__(self, model, domain, gridsize=10000, num_samples=10):
        super(MinValueEntropySearch, self).__init__(model)
        assert self.data[1].shape[1] == 1
        self.gridsize = gridsize
        self.num_samples = num_samples
        self.samples = np.zeros(num_samples)
        self._domain = domain

    def _setup(self):
        super(MinValueEntropySearch, self)._setup()
        model = self.models[0]
        X = self.data[0][self.feasible_data_index(), :]
        n_data = X.shape[0]
        grid = RandomDesign(self.gridsize, self._domain).generate()
        fmean, fvar = model.predict_f(np.vstack((X, grid)))
        fstd = np.sqrt(fvar)

        def prob_min_above(y):
            # P(every f(x) on data + grid exceeds y) under independent marginals
            return np.exp(np.sum(norm.logcdf(-(y - fmean) / fstd), axis=0))

        # bracket: right end = best posterior mean at the data; left end pushed down until P(min > left) >= 0.75
        right = fmean[np.argmin(fmean[:n_data])].flatten()
        left = right
        floor = np.min(fmean - 5.0 * fstd)
        step = 0
        while prob_min_above(left) < 0.75:
            w = 2.0 ** step
            left = w * floor + (1.0 - w) * right
            step += 1

        quartiles = [bisect(lambda y, q=q: prob_min_above(y) - q, left, right, maxiter=10000, xtol=0.01)
                     for q in (0.25, 0.5, 0.75)]
        lower_q, median, upper_q = quartiles
        # Gumbel(alpha, beta) through the three quantiles, then num_samples draws of y*
        beta = (lower_q - upper_q) / (np.log(np.log(4.0 / 3.0)) - np.log(np.log(4.0)))
        alpha = median + beta * np.log(np.log(2.0))
        u = np.random.rand(self.num_samples)
        self.samples = np.asarray(-np.log(-np.log(u)) * beta + alpha, dtype=np.float64).ravel()

    def build_acquisition(self, Xcand):
        return Xcand.mes(self.models[0], self.samples)
