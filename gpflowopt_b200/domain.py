"""Optimisation domains: labelled box bounds, their concatenation, membership tests and the affine map between two boxes
(the public surface of gpflowopt/domain.py that the acquisition path and its optimisers use)."""
import numpy as np

from .transforms import LinearTransform


def _as_row_block(x, width):
    x = np.atleast_2d(x)
    assert x.ndim == 2 and x.shape[1] == width
    return x


class Domain(object):
    """An ordered collection of one-dimensional parameters; ``a + b`` concatenates, ``a >> b`` maps box a onto box b."""

    __hash__ = None

    def __init__(self, parameters):
        self._parameters = list(parameters)

    # ---- structure ---------------------------------------------------------------------------------------------------
    def __iter__(self):
        for parameter in self._parameters:
            for leaf in parameter:
                yield leaf

    def __add__(self, other):
        assert isinstance(other, Domain)
        return Domain(self._parameters + other._parameters)

    def __radd__(self, other):                     # sum() / np.sum start from 0
        return self if (isinstance(other, (int, float)) and other == 0) else NotImplemented

    def __getitem__(self, key):
        if isinstance(key, list):
            return np.sum([self[k] for k in key])
        if isinstance(key, str):
            key = [p.label for p in self._parameters].index(key)
        return self._parameters[key]

    def __eq__(self, other):
        return isinstance(other, Domain) and self._parameters == other._parameters

    def __ne__(self, other):
        return not (self == other)

    @property
    def size(self):
        return sum(p.size for p in self._parameters)

    # ---- bounds ------------------------------------------------------------------------------------------------------
    def _collect(self, attribute):
        return np.array([getattr(p, attribute) for p in self._parameters]).flatten()

    @property
    def lower(self):
        return self._collect('lower')

    @property
    def upper(self):
        return self._collect('upper')

    def __contains__(self, X):
        """Every row inside the box, bounds included up to ``np.isclose`` (predicate of gpflowopt/domain.py:68-73).  The
        tolerance test only runs on entries that fail the strict comparison -- none, for a design of 1e6 interior points."""
        X = np.atleast_2d(X)
        if X.shape[1] != self.size:
            return False
        for bound, outside in ((self.lower, ~(self.lower < X)), (self.upper, ~(X < self.upper))):
            if outside.any() and not np.all(np.isclose(np.broadcast_to(bound, X.shape)[outside], X[outside])):
                return False
        return True

    def __rshift__(self, other):
        """The diagonal affine map taking this box onto ``other``: x -> A x + b with A the ratio of the widths and
        b = -upper * A + other.upper (the form of gpflowopt/domain.py:89-93, kept for bit-identical transforms)."""
        assert self.size == other.size
        A = (other.upper - other.lower) / (self.upper - self.lower)
        return LinearTransform(A, -self.upper * A + other.upper)

    # ---- current point -----------------------------------------------------------------------------------------------
    @property
    def value(self):
        return np.vstack([p.value for p in self._parameters]).T

    @value.setter
    def value(self, x):
        x = _as_row_block(x, self.size)
        widths = [p.size for p in self._parameters]
        for p, block in zip(self._parameters, np.split(x, np.cumsum(widths)[:-1], axis=1)):
            p.value = block


class Parameter(Domain):
    """A single labelled dimension; it is its own one-element domain."""

    size = 1

    def __init__(self, label, xinit):
        super(Parameter, self).__init__([self])
        self.label = label
        self._x = np.atleast_1d(xinit)

    def __iter__(self):
        yield self

    @property
    def value(self):
        return self._x

    @value.setter
    def value(self, x):
        self._x = np.atleast_1d(x).ravel()


class ContinuousParameter(Parameter):
    """A real interval [lb, ub]; the initial value defaults to its centre."""

    __hash__ = None

    def __init__(self, label, lb, ub, xinit=None):
        self._range = np.array([lb, ub], dtype=float)
        super(ContinuousParameter, self).__init__(label, (ub + lb) / 2.0 if xinit is None else xinit)

    def _bound(index):                                   # noqa: N805 -- property factory, not a method
        def getter(self):
            return self._range[index:index + 1].copy()

        def setter(self, value):
            self._range[index] = value
        return property(getter, setter)

    lower = _bound(0)
    upper = _bound(1)
    del _bound

    def __eq__(self, other):
        return isinstance(other, ContinuousParameter) and bool(np.array_equal(self._range, other._range))


class UnitCube(Domain):
    """[0, 1]^d"""

    def __init__(self, n_inputs):
        super(UnitCube, self).__init__([ContinuousParameter('u{0}'.format(i), 0, 1) for i in range(n_inputs)])
