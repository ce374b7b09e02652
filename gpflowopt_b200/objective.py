"""Objective wrapper on the optimiser call path (mirror of gpflowopt/objective.py:96-113 and of the GPflow
0.5.0 model.ObjectiveWrapper it derives from: non-finite gradients are replaced by zeros, last x is kept)."""
import numpy as np


class ObjectiveWrapper(object):
    def __init__(self, objective, exclude_gradient):
        self._objective = objective
        self._no_gradient = exclude_gradient
        self._previous_x = None
        self.counter = 0

    def __call__(self, x):
        x = np.atleast_2d(x)
        res = self._objective(x)
        if isinstance(res, tuple):
            f, g = res
            g_is_fin = np.isfinite(g)
            if np.all(g_is_fin):
                self._previous_x = x
            else:
                print("Warning: inf or nan in gradient: replacing with zeros")
                g = np.where(g_is_fin, g, 0.)
        else:                                  # gradient-free objectives may return f only
            f, g = res, None
            self._previous_x = x
        self.counter += x.shape[0]
        if self._no_gradient:
            return f
        return f, g
