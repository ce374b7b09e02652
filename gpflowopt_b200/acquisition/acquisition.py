"""Acquisition base classes: the drop-in surface of gpflowopt/acquisition/acquisition.py on top of libgpfo.

Kept from the reference: ``Acquisition(models, optimize_restarts)``, ``evaluate``, ``evaluate_with_gradients``,
``build_acquisition``, ``set_data``, ``enable_scaling``, ``models``, ``data``, ``feasible_data_index``,
``constraint_indices``, ``objective_indices``, ``_setup`` and the ``+`` / ``*`` operators, including the lazy
``_needs_setup`` protocol (constraints are set up before objectives).

Changed by design: ``build_acquisition(Xcand)`` receives a :class:`~gpflowopt_b200.engine.ProgramBuilder` instead
of a TensorFlow tensor and appends postfix ops; evaluation is one call into the CUDA library.
"""
import copy
from functools import wraps

import numpy as np

from .._lib import NotPositiveDefiniteError
from ..domain import UnitCube
from ..engine import Engine
from ..gpr import GPR
from ..scaling import DataScaler


def _best_restart(model, restarts):
    """First run from the current state, the others from ``model.randomize()``; failed runs are reported and skipped."""
    finished = []
    for attempt in range(1, restarts + 1):
        if attempt > 1:
            model.randomize()
        try:
            finished.append(model.optimize())
        except NotPositiveDefiniteError:                    # the reference catches tf.errors.InvalidArgumentError (Cholesky)
            print("Warning: optimization restart {0}/{1} failed".format(attempt, restarts))
    if not finished:
        raise RuntimeError("All model hyperparameter optimization restarts failed, exiting.")
    return min(finished, key=lambda run: run.fun)


def _join(kind, left, right):
    """``left (+|*) right`` with operands of the same aggregation kind flattened into one node."""
    flat = []
    for node in (left, right):
        flat.extend(node.operands if isinstance(node, kind) else [node])
    return kind(flat)


def setup_required(method):
    """Runs model optimisation and ``_setup`` of the whole tree first if the root is flagged
    (gpflowopt/acquisition/acquisition.py:32-54)."""
    @wraps(method)
    def runnable(instance, *args, **kwargs):
        assert isinstance(instance, Acquisition)
        hp = instance.highest_parent
        if hp._needs_setup:
            hp._needs_setup = False          # avoids recursion through feasible_data_index -> evaluate
            hp._optimize_models()
            hp._setup()
        return method(instance, *args, **kwargs)
    return runnable


class Acquisition(object):
    def __init__(self, models=[], optimize_restarts=5):
        models = list(np.atleast_1d(models))
        assert all(isinstance(m, (GPR, DataScaler)) for m in models), \
            "only GP regression models are supported by the device path"
        self._models = [m if isinstance(m, DataScaler) else DataScaler(m) for m in models]
        assert optimize_restarts >= 0
        self.optimize_restarts = optimize_restarts
        self._parent = None
        self._engine = None
        self._needs_setup = True

    # ---- tree plumbing -----------------------------------------------------------------------------------
    @property
    def highest_parent(self):
        node = self
        while node._parent is not None:
            node = node._parent
        return node

    def _adopt(self, child):
        child._parent = self
        self.highest_parent._needs_setup = True

    def _all_models(self):
        """Every model the device needs for this (sub)tree.  Equals ``models`` except below an MCMCAcquistion, whose
        ``models`` -- as in the reference -- shows the first copy only."""
        return self.models

    def _get_engine(self):
        root = self.highest_parent
        if root._engine is None or root._engine.ctx is None:
            adopted = [m._engine for m in root._all_models() if getattr(m, '_engine', None) is not None
                       and m._engine.ctx is not None]
            root._engine = adopted[0] if adopted else Engine()
        return root._engine

    # ---- reference API -----------------------------------------------------------------------------------
    def _optimize_models(self):
        """Refits every model ``optimize_restarts`` times and keeps the best run of each (protocol of
        gpflowopt/acquisition/acquisition.py:92-122; the fit itself is the model's business -- ``GPR.optimize`` keeps the
        given hyper-parameters)."""
        for model in (self.models if self.optimize_restarts > 0 else ()):
            model.set_state(_best_restart(model, self.optimize_restarts).x)

    def build_acquisition(self, Xcand):
        raise NotImplementedError

    def enable_scaling(self, domain):
        """Inputs of every model are mapped from ``domain`` to the unit cube and outputs are standardised."""
        d = self.data[0].shape[1]
        assert d == domain.size
        to_cube = domain >> UnitCube(d)
        for scaler in self.models:
            scaler.input_transform = to_cube
            scaler.normalize_output = True
        self.highest_parent._needs_setup = True

    def set_data(self, X, Y):
        """New training data for all models; model k takes its own columns of Y.  Returns the number of columns used."""
        widths = [scaler.Y.shape[1] for scaler in self.models]
        ends = np.cumsum(widths)
        for scaler, end, width in zip(self.models, ends, widths):
            scaler.X = X
            scaler.Y = Y[:, end - width:end]
        self.highest_parent._needs_setup = True
        return int(ends[-1]) if widths else 0

    @property
    def models(self):
        return list(self._models)

    @property
    def data(self):
        return self.models[0].X.value, np.hstack([m.Y.value for m in self.models])

    def constraint_indices(self):
        return np.empty((0,), dtype=int)

    def objective_indices(self):
        return np.setdiff1d(np.arange(self.data[1].shape[1]), self.constraint_indices())

    def feasible_data_index(self):
        return np.ones(self.data[0].shape[0], dtype=bool)

    def _setup(self):
        pass

    def _setup_constraints(self):
        if self.constraint_indices().size > 0:
            self._setup()

    def _setup_objectives(self):
        if self.constraint_indices().size == 0:
            self._setup()

    @setup_required
    def evaluate_with_gradients(self, Xcand):
        """Scores (N x 1) and their gradients w.r.t. the candidates (N x D)."""
        return self._get_engine().evaluate(self, Xcand, gradients=True)

    @setup_required
    def evaluate(self, Xcand):
        """Scores (N x 1)."""
        return self._get_engine().evaluate(self, Xcand, gradients=False)

    @setup_required
    def argmax(self, Xcand):
        """(best value, lowest index attaining it) over the candidate rows; scores stay on the device.
        This is the fused form of ``np.argmin(-evaluate(X))`` in MCOptimizer._optimize (gpflowopt/optim.py:155-164)."""
        return self._get_engine().argmax(self, Xcand)

    @setup_required
    def argmax_random(self, seed, first_row, n, lower, upper):
        """``argmax`` over candidates generated on the device (rows first_row .. first_row+n-1 of the stream keyed by
        ``seed``, uniform in the box [lower, upper]); nothing of size N crosses PCIe in either direction."""
        return self._get_engine().argmax_random(self, seed, first_row, n, lower, upper)

    @setup_required
    def evaluate_random(self, seed, first_row, n, lower, upper):
        """``evaluate`` over device-generated candidates: only the n scores come back (``_lib.random_rows`` regenerates rows)."""
        return self._get_engine().evaluate_random(self, seed, first_row, n, lower, upper)

    def __add__(self, other):
        return _join(AcquisitionSum, self, other)

    def __mul__(self, other):
        return _join(AcquisitionProduct, self, other)


class AcquisitionAggregation(Acquisition):
    """Reduces the scores of several acquisitions (gpflowopt/acquisition/acquisition.py:301-365)."""

    def __init__(self, operands, oper):
        super(AcquisitionAggregation, self).__init__()
        assert all(isinstance(x, Acquisition) for x in operands)
        assert oper in ('sum', 'prod')
        self.operands = list(operands)
        self._oper = oper
        for o in self.operands:
            self._adopt(o)

    def _optimize_models(self):
        for oper in self.operands:
            oper._optimize_models()

    @property
    def models(self):
        return [model for acq in self.operands for model in acq.models]

    def _all_models(self):
        return [model for acq in self.operands for model in acq._all_models()]

    def enable_scaling(self, domain):
        for oper in self.operands:
            oper.enable_scaling(domain)

    def set_data(self, X, Y):
        used = 0
        for node in self.operands:              # every operand consumes its columns from the left
            used += node.set_data(X, Y[:, used:])
        return used

    def _setup_constraints(self):
        for node in self.operands:
            if node.constraint_indices().size:
                node._setup_constraints()

    def _setup_objectives(self):
        for node in self.operands:
            node._setup_objectives()

    def _setup(self):
        # constraint acquisitions first: the incumbent of an objective acquisition is taken over the points they accept
        self._setup_constraints()
        self._setup_objectives()

    def constraint_indices(self):
        """Constraint columns of the concatenated output matrix.  As in the reference (acquisition.py:349-355, pinned by
        tests/golden/tree_golden.npz) operand k's local indices are shifted by the width of operand k-1 -- equal to the
        cumulative offset for the usual two-operand trees, and kept as is for three or more so that results stay identical."""
        widths = [node.data[1].shape[1] for node in self.operands]
        shifts = [0] + widths[:-1]
        return np.concatenate([node.constraint_indices() + shift for node, shift in zip(self.operands, shifts)]).astype(int)

    def feasible_data_index(self):
        return np.all(np.vstack([o.feasible_data_index() for o in self.operands]), axis=0)

    def build_acquisition(self, Xcand):
        # folded left to right (a b + c + ...): the same summation order as one k-ary reduce, but the program's stack stays
        # two deep whatever the number of operands (the device stack holds 8 values)
        node = self.operands[0].build_acquisition(Xcand)
        for operand in self.operands[1:]:
            node = Xcand.reduce(self._oper, [node, operand.build_acquisition(Xcand)])
        return node

    def __getitem__(self, item):
        return self.operands[item]


class AcquisitionSum(AcquisitionAggregation):
    def __init__(self, operands):
        super(AcquisitionSum, self).__init__(operands, 'sum')


class AcquisitionProduct(AcquisitionAggregation):
    def __init__(self, operands):
        super(AcquisitionProduct, self).__init__(operands, 'prod')


class MCMCAcquistion(AcquisitionSum):
    """Mean of an acquisition over hyper-parameter draws (gpflowopt/acquisition/acquisition.py:396-458).

    Drawing the samples (GPflow's HMC) is outside this path: pass them as ``hyper_draws`` (n_slices x n_hypers
    free-state rows per model, see ``GPR.get_free_state``); the evaluation = SUM over copies then SCALE 1/n on
    the device."""

    def __init__(self, acquisition, n_slices, hyper_draws=None, **kwargs):
        assert isinstance(acquisition, Acquisition)
        assert n_slices > 0
        copies = [acquisition] + [copy.deepcopy(acquisition) for _ in range(n_slices - 1)]
        for c in copies[1:]:
            c.optimize_restarts = 0
            c._engine = None
            for m in c._models:
                m._engine = None
        super(MCMCAcquistion, self).__init__(copies)
        self._sample_opt = kwargs
        if hyper_draws is not None:
            self.set_hyper_draws(hyper_draws)

    def set_hyper_draws(self, hyper_draws):
        hyper_draws = np.atleast_2d(hyper_draws)
        assert hyper_draws.shape[0] == len(self.operands)
        for draw, oper in zip(hyper_draws, self.operands):
            offset = 0
            for m in oper.models:
                n = m.get_free_state().size
                m.set_state(draw[offset:offset + n])
                offset += n
        self.highest_parent._needs_setup = True

    def _optimize_models(self):
        self.operands[0]._optimize_models()

    @property
    def models(self):
        """Only the models of the first operand; the copies stay hidden (gpflowopt/acquisition/acquisition.py:434-437), so
        ``data`` and the objective / constraint indices keep the width of ONE acquisition."""
        return self.operands[0].models

    def _all_models(self):
        return [model for acq in self.operands for model in acq._all_models()]

    def enable_scaling(self, domain):
        for operand in self.operands:
            operand.enable_scaling(domain)

    def constraint_indices(self):
        return self.operands[0].constraint_indices()

    def feasible_data_index(self):
        return self.operands[0].feasible_data_index()

    def set_data(self, X, Y):
        for operand in self.operands:
            offset = operand.set_data(X, Y)
        return offset

    def build_acquisition(self, Xcand):
        node = super(MCMCAcquistion, self).build_acquisition(Xcand)
        return Xcand.scale(node, 1. / len(self.operands))
