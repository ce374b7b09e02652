"""Host-side mirror of the reference's optimisation domain (gpflowopt/domain.py): box bounds -> affine maps."""
from itertools import chain

import numpy as np

from .transforms import LinearTransform


class Domain(object):
    """A (product) space over which is optimised (gpflowopt/domain.py:22-129)."""

    def __init__(self, parameters):
        self._parameters = list(parameters)

    @property
    def lower(self):
        return np.array([p.lower for p in self._parameters]).flatten()

    @property
    def upper(self):
        return np.array([p.upper for p in self._parameters]).flatten()

    @property
    def size(self):
        return sum(p.size for p in self._parameters)

    def __add__(self, other):
        assert isinstance(other, Domain)
        return Domain(self._parameters + other._parameters)

    def __radd__(self, other):          # lets np.sum / sum() build a domain from parameters
        if isinstance(other, (int, float)) and other == 0:
            return self
        return NotImplemented

    def __eq__(self, other):
        return isinstance(other, Domain) and self._parameters == other._parameters

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    def __contains__(self, X):
        X = np.atleast_2d(X)
        if X.shape[1] != self.size:
            return False
        # same predicate as the reference (domain.py:68-73): inside, or np.isclose to a bound -- with the tolerance test run
        # only on the entries that fail the strict comparison (none, for a design of 1e6 interior points)
        for bound, outside in ((self.lower, ~(self.lower < X)), (self.upper, ~(X < self.upper))):
            if outside.any() and not np.all(np.isclose(np.broadcast_to(bound, X.shape)[outside], X[outside])):
                return False
        return True

    def __iter__(self):
        for v in chain(*map(iter, self._parameters)):
            yield v

    def __getitem__(self, items):
        if isinstance(items, list):
            return np.sum([self[item] for item in items])
        if isinstance(items, str):
            items = [p.label for p in self._parameters].index(items)
        return self._parameters[items]

    def __rshift__(self, other):
        """Affine map of this box onto ``other`` (gpflowopt/domain.py:89-93)."""
        assert self.size == other.size
        A = (other.upper - other.lower) / (self.upper - self.lower)
        b = -self.upper * A + other.upper
        return LinearTransform(A, b)

    @property
    def value(self):
        return np.vstack([p.value for p in self._parameters]).T

    @value.setter
    def value(self, x):
        x = np.atleast_2d(x)
        assert x.ndim == 2 and x.shape[1] == self.size
        offset = 0
        for p in self._parameters:
            p.value = x[:, offset:offset + p.size]
            offset += p.size


class Parameter(Domain):
    def __init__(self, label, xinit):
        super(Parameter, self).__init__([self])
        self.label = label
        self._x = np.atleast_1d(xinit)

    @property
    def size(self):
        return 1

    def __iter__(self):
        yield self

    @property
    def value(self):
        return self._x

    @value.setter
    def value(self, x):
        self._x = np.atleast_1d(x).ravel()


class ContinuousParameter(Parameter):
    def __init__(self, label, lb, ub, xinit=None):
        self._range = np.array([lb, ub], dtype=float)
        super(ContinuousParameter, self).__init__(label, xinit if xinit is not None else (ub + lb) / 2.0)

    @property
    def lower(self):
        return np.array([self._range[0]])

    @lower.setter
    def lower(self, value):
        self._range[0] = value

    @property
    def upper(self):
        return np.array([self._range[1]])

    @upper.setter
    def upper(self, value):
        self._range[1] = value

    def __eq__(self, other):
        return isinstance(other, ContinuousParameter) and bool(self.lower == other.lower) and \
            bool(self.upper == other.upper)

    __hash__ = None


class UnitCube(Domain):
    """[0, 1]^d"""

    def __init__(self, n_inputs):
        super(UnitCube, self).__init__([ContinuousParameter('u{0}'.format(i), 0, 1) for i in range(n_inputs)])
