"""Minimal GP-regression parameter holders: the attribute surface of GPflow 0.5.0 that GPflowOpt touches
(X, Y, kern.variance, kern.lengthscales, likelihood.variance, predict_f).

These classes hold numbers only.  All arithmetic (kernel matrix, Cholesky, prediction) runs on the device
through libgpfo; hyper-parameter learning (GPflow's Model.optimize / HMC) is outside this path, so
``optimize()`` keeps the hyper-parameters it was given.
"""
import warnings

import numpy as np
from scipy.optimize import OptimizeResult

from . import _lib


class DataHolder(object):
    """``.value`` / ``.shape`` view of an array, as GPflow's DataHolder offers."""

    def __init__(self, array):
        self._array = np.asarray(array, dtype=np.float64)

    @property
    def value(self):
        return self._array.copy()

    @property
    def shape(self):
        return self._array.shape


def _same(a, b):
    if isinstance(a, DataHolder) and isinstance(b, DataHolder):
        return a.shape == b.shape and np.array_equal(a._array, b._array)
    if isinstance(a, (float, int, np.ndarray)) and isinstance(b, (float, int, np.ndarray)):
        return np.shape(a) == np.shape(b) and bool(np.array_equal(a, b))
    return a is b


class _Versioned(object):
    """Bumps ``version`` on every public attribute assignment so engines can see stale device state."""

    def __setattr__(self, key, value):
        old = self.__dict__.get(key, _Versioned)
        object.__setattr__(self, key, value)
        if key.startswith('_') or key == 'version':
            return
        if old is not _Versioned and _same(old, value):
            return                       # e.g. set_state() with unchanged hyper-parameters: no re-factorisation
        object.__setattr__(self, 'version', getattr(self, 'version', 0) + 1)


class Stationary(_Versioned):
    kind = None

    def __init__(self, input_dim, variance=1.0, lengthscales=None, ARD=False):
        self.input_dim = int(input_dim)
        self.ARD = bool(ARD)
        self.variance = float(variance)
        if lengthscales is None:
            lengthscales = np.ones(self.input_dim) if ARD else 1.0
        self.lengthscales = lengthscales

    def __setattr__(self, key, value):
        if key == 'lengthscales':
            value = np.atleast_1d(np.asarray(value, dtype=np.float64)).copy()
        super(Stationary, self).__setattr__(key, value)


class RBF(Stationary):
    kind = _lib.KERN_RBF


class Matern52(Stationary):
    kind = _lib.KERN_MATERN52


class Matern32(Stationary):
    kind = _lib.KERN_MATERN32


class Gaussian(_Versioned):
    def __init__(self, variance=1.0):
        self.variance = float(variance)


class GPR(_Versioned):
    """Gaussian process regression with a Gaussian likelihood and a zero mean function."""

    def __init__(self, X, Y, kern, noise_variance=1.0, name='GPR'):
        self.kern = kern
        self.likelihood = Gaussian(noise_variance)
        self.name = name
        self.X = X
        self.Y = Y

    def __setattr__(self, key, value):
        if key in ('X', 'Y'):
            value = value if isinstance(value, DataHolder) else DataHolder(np.atleast_2d(value))
            assert len(value.shape) == 2
        super(GPR, self).__setattr__(key, value)

    @property
    def state_version(self):
        return (self.version, self.kern.version, self.likelihood.version)

    # Hyper-parameter learning (GPflow's Model.optimize / HMC) is not part of this path (SURVEY.md section 2, rows 1 and 14;
    # out of scope per section 8).  optimize() keeps the hyper-parameters it was given and says so once per process: a
    # BO loop built on these holders scores candidates under FIXED hyper-parameters unless the caller assigns new ones
    # (kern.variance / kern.lengthscales / likelihood.variance, or set_state) between iterations.
    _warned_fixed = False

    def optimize(self, **kwargs):
        if not GPR._warned_fixed:
            GPR._warned_fixed = True
            warnings.warn("gpflowopt_b200.GPR.optimize() does not fit hyper-parameters (model training is outside the "
                          "accelerated path); the values assigned to the model are kept.  Assign fitted values yourself or "
                          "construct the acquisition with optimize_restarts=0 to silence this.", UserWarning, stacklevel=2)
        return OptimizeResult(x=self.get_free_state(), fun=0.0, success=True, message="hyper-parameters kept")

    def randomize(self):
        pass

    def get_free_state(self):
        return np.hstack([[self.kern.variance], self.kern.lengthscales, [self.likelihood.variance]])

    def set_state(self, x):
        x = np.asarray(x, dtype=np.float64).ravel()
        self.kern.variance = float(x[0])
        self.kern.lengthscales = x[1:-1]
        self.likelihood.variance = float(x[-1])
