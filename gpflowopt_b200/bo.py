"""BayesianOptimizer: the caller of the hot path (mirror of gpflowopt/bo.py), with the caller-side fix-ups of
SURVEY.md section 8(f) rank 1:

* the acquisition objective handed to the optimiser carries ``.acquisition`` so that Monte-Carlo / candidate stages
  run the fused device argmax (no N-sized device->host copy), and it only asks for gradients when the receiving
  optimiser is gradient based (the reference always builds the gradient graph, bo.py:244-245, and drops it later);
* the GP factorisation lives on the device across the MC stage and all L-BFGS-B polish steps of one iteration
  (the reference re-runs the Cholesky inside every session.run).

Cholesky failures surface as ``NotPositiveDefiniteError`` (the reference's tf.errors.InvalidArgumentError) and are
handled by ``jitchol_callback`` exactly as there: inflate the likelihood variance until the factorisation succeeds.
"""
import os
from contextlib import contextmanager

import numpy as np
from scipy.optimize import OptimizeResult

from ._lib import NotPositiveDefiniteError
from .acquisition import Acquisition, MCMCAcquistion
from .design import Design, EmptyDesign
from .gpr import GPR
from .objective import ObjectiveWrapper
from .optim import Optimizer, SciPyOptimizer
from .pareto import non_dominated_sort
from .scaling import DataScaler


def jitchol_callback(models):
    """Increase the likelihood variance in case of Cholesky failures (gpflowopt/bo.py:30-52).

    ``models`` are GPRs or the DataScalers around them.  Handing over the acquisition's own DataScalers (what
    BayesianOptimizer does) makes the probe factorisation the one the following evaluation re-uses: same engine, same
    device state, nothing is factorised twice."""
    for m in np.atleast_1d(models):
        scaler = m if isinstance(m, DataScaler) else None
        gpr = m.wrapped if scaler is not None else m
        if not isinstance(gpr, GPR):
            continue
        base = gpr.likelihood.variance
        eKdiag = gpr.kern.variance                         # mean(diag(K_symm)) of a stationary kernel, no noise (bo.py:45)
        probe = scaler if scaler is not None else DataScaler(gpr)
        for e in [0] + [10 ** ex for ex in range(-6, -1)]:
            try:
                gpr.likelihood.variance = base + e * eKdiag
                probe.predict_f(probe.X.value[:1])         # forces a factorisation
                break
            except NotPositiveDefiniteError:               # pragma: no cover
                gpr.likelihood.variance = base


class _InverseAcquisition(object):
    """x -> -acquisition(x) (and -gradient when asked for); carries the acquisition for the fused MC path."""

    def __init__(self, acquisition, with_gradient):
        self.acquisition = acquisition
        self.with_gradient = with_gradient

    def bind_gradient(self, with_gradient):
        """The objective as one optimiser stage wants it (called by Optimizer.optimize): gradient-based stages get
        (-value, -gradient), gradient-free ones the value-only kernel.  Inside a StagedOptimizer every stage binds its own."""
        return self if with_gradient == self.with_gradient else _InverseAcquisition(self.acquisition, with_gradient)

    def __call__(self, x):
        x = np.atleast_2d(x)
        if self.with_gradient:
            return tuple(-r for r in self.acquisition.evaluate_with_gradients(x))
        return -self.acquisition.evaluate(x), np.zeros((x.shape[0], 0))


def _summarise(acquisition, success, message):
    """What a run reports (reference behaviour, gpflowopt/bo.py:156-203): among the evaluated points that satisfy every
    constraint model -- or all of them, flagged unsuccessful, if none does -- the single best one (one objective), the
    non-dominated set (several objectives) or the whole feasible set (constraints only)."""
    X, Y = acquisition.data
    feasible = acquisition.feasible_data_index()
    if feasible.any():
        X, Y = X[feasible], Y[feasible]
    else:
        success, message = False, "No evaluations satisfied all the constraints"
    objectives = Y[:, acquisition.objective_indices()]
    constraints = Y[:, acquisition.constraint_indices()]
    n_obj = objectives.shape[1]
    if n_obj == 0:
        keep = np.arange(constraints.shape[0])
    elif n_obj == 1:
        keep = np.argmin(objectives)
    else:
        keep = non_dominated_sort(objectives)[1] == 0
    return OptimizeResult(x=X[keep, :], fun=objectives[keep, :], constraints=constraints[keep, :], success=success,
                          message=message)


def _progress_line(iteration, summary, step):
    """One line per iteration in verbose mode: running minima of objectives and constraints, then any failure messages."""
    parts = []
    for label, block in (("fmin", summary.fun), ("constraints", summary.constraints)):
        if block.shape[0] > 0 and block.shape[1] > 0:
            lows = np.atleast_1d(np.min(block, axis=0))
            parts.append("%s [%s]" % (label, ", ".join("{:.3}".format(v) for v in lows)))
    parts.extend(r.message for r in (summary, step) if not r.success)
    return "iter #{0:>3} - {1}".format(iteration, " - ".join(str(p) for p in parts))


class BayesianOptimizer(Optimizer):
    """Sequential model-based optimisation loop (public surface of gpflowopt/bo.py:55-305): evaluate the initial design, then
    ``n_iter`` times { refresh the models, maximise the acquisition with ``optimizer``, evaluate the objectives there }."""

    def __init__(self, domain, acquisition, optimizer=None, initial=None, scaling=True, hyper_draws=None,
                 callback=jitchol_callback, verbose=False):
        for ok in (isinstance(acquisition, Acquisition), hyper_draws is None or hyper_draws > 0,
                   optimizer is None or isinstance(optimizer, Optimizer), initial is None or isinstance(initial, Design)):
            assert ok
        super(BayesianOptimizer, self).__init__(domain, exclude_gradient=True)
        self._scaling = scaling
        if scaling:
            acquisition.enable_scaling(domain)
        if hyper_draws is not None:
            acquisition = MCMCAcquistion(acquisition, hyper_draws)
        self.acquisition = acquisition
        self.optimizer = SciPyOptimizer(domain) if optimizer is None else optimizer
        self.optimizer.domain = domain
        self.set_initial((initial if initial is not None else EmptyDesign(domain)).generate())
        self._model_callback = callback
        self.verbose = verbose

    @property
    def domain(self):
        return self._domain

    @domain.setter
    def domain(self, dom):
        assert dom.size == self._domain.size
        self._domain = dom
        self.set_initial(dom.value)
        if self._scaling:
            self.acquisition.enable_scaling(dom)

    def _update_model_data(self, newX, newY):
        """Appends evaluations to the data every model of the acquisition is trained on."""
        oldX, oldY = self.acquisition.data
        assert newX.shape[-1] == oldX.shape[1] and newY.shape[-1] == oldY.shape[1] and newX.shape[0] == newY.shape[0]
        if newX.size > 0:
            self.acquisition.set_data(np.vstack((oldX, newX)), np.vstack((oldY, newY)))

    def _evaluate_objectives(self, X, fxs):
        """All objective / constraint callables at X, side by side; the (empty) second item is the gradient slot."""
        n_out = self.acquisition.data[1].shape[1]
        if X.size == 0:
            return np.empty((0, n_out)), np.zeros((0, 0))
        columns = np.hstack([f(X) for f in fxs])
        assert columns.shape[1] == n_out
        return columns, np.zeros((X.shape[0], 0))

    def _create_bo_result(self, success, message):
        return _summarise(self.acquisition, success, message)

    def optimize(self, objectivefx, n_iter=20):
        fxs = np.atleast_1d(objectivefx)
        return super(BayesianOptimizer, self).optimize(lambda x: self._evaluate_objectives(x, fxs), n_iter=n_iter)

    def _optimize(self, fx, n_iter):
        assert isinstance(fx, ObjectiveWrapper)
        start = self.get_initial()
        self._update_model_data(start, fx(start))
        self.set_initial(EmptyDesign(self.domain).generate())          # the initial design is consumed once
        # the optimiser minimises -acquisition; gradients are only asked for when it can use them
        target = _InverseAcquisition(self.acquisition, self.optimizer.gradient_enabled())
        iteration = 0
        while iteration < n_iter:
            with self.silent():
                if self._model_callback is not None and self.acquisition._needs_setup:
                    models = self.acquisition.models
                    # the default callback probes through the acquisition's own scalers (shared engine); a user callback
                    # receives the wrapped models, as in the reference (bo.py:250-251)
                    self._model_callback(models if self._model_callback is jitchol_callback else [m.wrapped for m in models])
                step = self.optimizer.optimize(target)
                self._update_model_data(step.x, fx(step.x))
            if self.verbose:
                with self.silent():
                    summary = self._create_bo_result(True, 'Monitor')
                print(_progress_line(iteration, summary, step))
            iteration += 1
        return self._create_bo_result(True, "OK")

    @contextmanager
    def failsafe(self):
        """Saves the evaluated data to failed_bopt_<id>.npz when the wrapped block raises (gpflowopt/bo.py:290-305)."""
        try:
            yield
        except Exception as exc:
            X, Y = self.acquisition.data
            np.savez('failed_bopt_{0}'.format(id(exc)), X=X, Y=Y)
            raise
