"""Candidate sources for the Monte-Carlo optimiser (mirror of gpflowopt/design.py, the parts on the path)."""
import numpy as np

from .domain import ContinuousParameter


_UNIT_CUBES = {}


class Design(object):
    """N points generated in a D-dimensional domain (gpflowopt/design.py:29-81)."""

    def __init__(self, size, domain):
        self.size = size
        self.domain = domain

    @property
    def generative_domain(self):
        # the unit cube of the domain's dimension; built once per dimension (a BO iteration's MCOptimizer(domain, 1000) call is
        # 0.2 ms in all, of which rebuilding this sum of parameters twice per generate() was 0.08 ms)
        d = self.domain.size
        if d not in _UNIT_CUBES:
            _UNIT_CUBES[d] = np.sum([ContinuousParameter('d{0}'.format(i), 0, 1) for i in range(d)])
        return _UNIT_CUBES[d]

    def generate(self):
        Xs = self.create_design()
        gen = self.generative_domain
        assert Xs in gen
        assert Xs.shape == (self.size, self.domain.size)
        X = (gen >> self.domain).forward(Xs)
        assert X in self.domain
        return X

    def create_design(self):
        raise NotImplementedError


class RandomDesign(Design):
    """Uniform random points: the MC candidate source (gpflowopt/design.py:83-94)."""

    def create_design(self):
        return np.asarray(np.random.rand(self.size, self.domain.size), dtype=np.float64)


class FactorialDesign(Design):
    """k-level grid (gpflowopt/design.py:97-116)."""

    def __init__(self, levels, domain):
        self.levels = levels
        super(FactorialDesign, self).__init__(levels ** domain.size, domain)

    @property
    def generative_domain(self):
        return self.domain

    def create_design(self):
        Xs = np.meshgrid(*[np.linspace(l, u, self.levels) for l, u in zip(self.domain.lower, self.domain.upper)])
        return np.vstack([X.ravel() for X in Xs]).T


class EmptyDesign(Design):
    def __init__(self, domain):
        super(EmptyDesign, self).__init__(0, domain)

    def create_design(self):
        return np.empty((0, self.domain.size))
