"""Single-model acquisition leaves: Expected Improvement, Probability of Improvement, Probability of Feasibility and the
Lower Confidence Bound (public surface of gpflowopt/acquisition/{ei,poi,pof,lcb}.py).

Each class contributes one postfix op to the device program; the arithmetic itself (ei.py:70-79, poi.py:63-67,
pof.py:81-85, lcb.py:54-57) lives in csrc/gpfo_internal.cuh::gpfo_run_program."""
import numpy as np

from .acquisition import Acquisition


class _IncumbentLeaf(Acquisition):
    """Leaves that compare against the incumbent: the lowest posterior mean over the training inputs every constraint
    model of the tree accepts (ei.py:63-68, poi.py:57-61)."""

    _op = None                                   # name of the ProgramBuilder method

    def __init__(self, model):
        super(_IncumbentLeaf, self).__init__(model)
        self.fmin = np.zeros(1)
        self._setup()

    def _setup(self):
        super(_IncumbentLeaf, self)._setup()
        accepted = self.highest_parent.feasible_data_index()
        means, _ = self.models[0].predict_f(self.data[0][accepted, :])
        self.fmin = means.min(axis=0)

    def build_acquisition(self, Xcand):
        return getattr(Xcand, self._op)(self.models[0], self.fmin)


class ExpectedImprovement(_IncumbentLeaf):
    _op = 'ei'


class ProbabilityOfImprovement(_IncumbentLeaf):
    _op = 'poi'


class ProbabilityOfFeasibility(Acquisition):
    """P(constraint value < threshold).  Every output of its model is a constraint, and a training point counts as feasible
    when this probability exceeds ``minimum_pof`` there (pof.py:63-79)."""

    def __init__(self, model, threshold=0.0, minimum_pof=0.5):
        super(ProbabilityOfFeasibility, self).__init__(model)
        self.threshold, self.minimum_pof = threshold, minimum_pof

    def constraint_indices(self):
        return np.arange(self.data[1].shape[1])

    def feasible_data_index(self):
        return self.evaluate(self.data[0]).ravel() > self.minimum_pof

    def build_acquisition(self, Xcand):
        return Xcand.pof(self.models[0], self.threshold)


class LowerConfidenceBound(Acquisition):
    """mean - sigma * sqrt(variance), to be minimised through the usual sign flip (lcb.py:54-57)."""

    def __init__(self, model, sigma=2.0):
        super(LowerConfidenceBound, self).__init__(model)
        self.sigma = np.array(sigma, dtype=np.float64)

    def build_acquisition(self, Xcand):
        return Xcand.lcb(self.models[0], self.sigma)
