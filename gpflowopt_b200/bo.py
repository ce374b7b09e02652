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
    """Increase the likelihood variance in case of Cholesky failures (gpflowopt/bo.py:30-52)."""
    for m in np.atleast_1d(models):
        scaler = m if isinstance(m, DataScaler) else None
        gpr = m.wrapped if scaler is not None else m
        if not isinstance(gpr, GPR):
            continue
        base = gpr.likelihood.variance
        eKdiag = gpr.kern.variance + base                  # mean diagonal of K for a stationary kernel
        probe = scaler if scaler is not None else DataScaler(gpr)
        for e in [0] + [10 ** ex for ex in range(-6, -1)]:
            try:
                gpr.likelihood.variance = base + e * eKdiag
                probe.predict_f(gpr.X.value[:1] if scaler is None else scaler.X.value[:1])   # forces a factorisation
                break
            except NotPositiveDefiniteError:               # pragma: no cover
                gpr.likelihood.variance = base


class _InverseAcquisition(object):
    """x -> -acquisition(x) (and -gradient when asked for); carries the acquisition for the fused MC path."""

    def __init__(self, acquisition, with_gradient):
        self.acquisition = acquisition
        self.with_gradient = with_gradient

    def __call__(self, x):
        x = np.atleast_2d(x)
        if self.with_gradient:
            return tuple(-r for r in self.acquisition.evaluate_with_gradients(x))
        return -self.acquisition.evaluate(x), np.zeros((x.shape[0], 0))


class BayesianOptimizer(Optimizer):
    def __init__(self, domain, acquisition, optimizer=None, initial=None, scaling=True, hyper_draws=None,
                 callback=jitchol_callback, verbose=False):
        assert isinstance(acquisition, Acquisition)
        assert hyper_draws is None or hyper_draws > 0
        assert optimizer is None or isinstance(optimizer, Optimizer)
        assert initial is None or isinstance(initial, Design)
        super(BayesianOptimizer, self).__init__(domain, exclude_gradient=True)
        self._scaling = scaling
        if self._scaling:
            acquisition.enable_scaling(domain)
        self.acquisition = acquisition if hyper_draws is None else MCMCAcquistion(acquisition, hyper_draws)
        self.optimizer = optimizer or SciPyOptimizer(domain)
        self.optimizer.domain = domain
        initial = initial or EmptyDesign(domain)
        self.set_initial(initial.generate())
        self._model_callback = callback
        self.verbose = verbose

    @property
    def domain(self):
        return self._domain

    @domain.setter
    def domain(self, dom):
        assert self._domain.size == dom.size
        self._domain = dom
        self.set_initial(dom.value)
        if self._scaling:
            self.acquisition.enable_scaling(dom)

    def _update_model_data(self, newX, newY):
        assert self.acquisition.data[0].shape[1] == newX.shape[-1]
        assert self.acquisition.data[1].shape[1] == newY.shape[-1]
        assert newX.shape[0] == newY.shape[0]
        if newX.size == 0:
            return
        X = np.vstack((self.acquisition.data[0], newX))
        Y = np.vstack((self.acquisition.data[1], newY))
        self.acquisition.set_data(X, Y)

    def _evaluate_objectives(self, X, fxs):
        if X.size > 0:
            evaluations = np.hstack([f(X) for f in fxs])
            assert evaluations.shape[1] == self.acquisition.data[1].shape[1]
            return evaluations, np.zeros((X.shape[0], 0))
        return np.empty((0, self.acquisition.data[1].shape[1])), np.zeros((0, 0))

    def _create_bo_result(self, success, message):
        """Best feasible point / Pareto set / feasible set, as in gpflowopt/bo.py:156-203."""
        X, Y = self.acquisition.data
        valid = self.acquisition.feasible_data_index()
        if np.any(valid):
            X, Y = X[valid, :], Y[valid, :]
        else:
            success = False
            message = "No evaluations satisfied all the constraints"
        Yo = Y[:, self.acquisition.objective_indices()]
        Yc = Y[:, self.acquisition.constraint_indices()]
        if Yo.shape[1] == 1:
            idx = np.argmin(Yo)
        elif Yo.shape[1] > 1:
            _, dom = non_dominated_sort(Yo)
            idx = dom == 0
        else:
            idx = np.arange(Yc.shape[0])
        return OptimizeResult(x=X[idx, :], success=success, fun=Yo[idx, :], constraints=Yc[idx, :], message=message)

    def optimize(self, objectivefx, n_iter=20):
        fxs = np.atleast_1d(objectivefx)
        return super(BayesianOptimizer, self).optimize(lambda x: self._evaluate_objectives(x, fxs), n_iter=n_iter)

    def _optimize(self, fx, n_iter):
        assert isinstance(fx, ObjectiveWrapper)
        initial = self.get_initial()
        values = fx(initial)
        self._update_model_data(initial, values)
        self.set_initial(EmptyDesign(self.domain).generate())
        inverse_acquisition = _InverseAcquisition(self.acquisition, self.optimizer.gradient_enabled())
        for i in range(n_iter):
            with self.silent():
                if self._model_callback and self.acquisition._needs_setup:
                    self._model_callback([m.wrapped for m in self.acquisition.models])
                result = self.optimizer.optimize(inverse_acquisition)
                self._update_model_data(result.x, fx(result.x))
            if self.verbose:
                with self.silent():
                    bo_result = self._create_bo_result(True, 'Monitor')
                metrics = []
                if bo_result.fun.shape[0] > 0:
                    funs = np.atleast_1d(np.min(bo_result.fun, axis=0))
                    metrics.append('fmin [' + ', '.join('{:.3}'.format(f) for f in funs) + ']')
                if bo_result.constraints.shape[0] > 0 and bo_result.constraints.shape[1] > 0:
                    cons = np.atleast_1d(np.min(bo_result.constraints, axis=0))
                    metrics.append('constraints [' + ', '.join('{:.3}'.format(c) for c in cons) + ']')
                metrics += [r.message for r in [bo_result, result] if not r.success]
                print('iter #{0:>3} - {1}'.format(i, ' - '.join(str(m) for m in metrics)))
        return self._create_bo_result(True, "OK")

    @contextmanager
    def failsafe(self):
        """Dumps the evaluated data to failed_bopt_<id>.npz when the wrapped block raises (gpflowopt/bo.py:290-305)."""
        try:
            yield
        except Exception as e:
            X, Y = self.acquisition.data
            np.savez('failed_bopt_{0}'.format(id(e)), X=X, Y=Y)
            raise
