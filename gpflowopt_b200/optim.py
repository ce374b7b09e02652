"""Acquisition optimisers: mirror of gpflowopt/optim.py (GPflowOpt 0.1.1, Apache-2.0; the Optimizer / MCOptimizer /
SciPyOptimizer / StagedOptimizer surface and messages follow it so results stay identical, see tests/golden/optim_golden.npz) with the candidate scoring + argmin of
MCOptimizer / CandidateOptimizer fused on the device when the objective is a gpflowopt_b200 acquisition.

An objective handed to ``Optimizer.optimize`` may carry an ``acquisition`` attribute (see
``bo.inverse_acquisition``).  For such objectives the Monte-Carlo stage calls ``Acquisition.argmax`` -- one
gpfo_eval with a block argmax and no N-sized device->host copy -- and, under torch.distributed, shards the
candidate rows over the ranks (``gpflowopt_b200.dist``).  Any other callable goes through the generic
evaluate-then-np.argmin route of the reference.
"""
import contextlib
import os
import sys
import threading
import warnings

import numpy as np
from scipy.optimize import OptimizeResult, minimize

from . import dist
from .design import RandomDesign
from .objective import ObjectiveWrapper


class Optimizer(object):
    def __init__(self, domain, exclude_gradient=False):
        self._domain = domain
        self._initial = domain.value
        self._wrapper_args = dict(exclude_gradient=exclude_gradient)

    @property
    def domain(self):
        return self._domain

    @domain.setter
    def domain(self, dom):
        self._domain = dom
        self.set_initial(dom.value)

    def optimize(self, objectivefx, **kwargs):
        """Minimises ``objectivefx`` over the domain; Ctrl-C returns the last good point (optim.py:63-85)."""
        if hasattr(objectivefx, 'bind_gradient'):          # acquisition objectives adapt to what this stage consumes
            objectivefx = objectivefx.bind_gradient(self.gradient_enabled())
        objective = ObjectiveWrapper(objectivefx, **self._wrapper_args)
        objective.acquisition = getattr(objectivefx, 'acquisition', None)
        try:
            result = self._optimize(objective, **kwargs)
        except KeyboardInterrupt:
            result = OptimizeResult(x=objective._previous_x, success=False,
                                    message="Caught KeyboardInterrupt, returning last good value.")
        result.x = np.atleast_2d(result.x)
        result.nfev = objective.counter
        return result

    def get_initial(self):
        return self._initial

    def set_initial(self, initial):
        initial = np.atleast_2d(initial)
        assert initial in self.domain
        self._initial = initial

    def gradient_enabled(self):
        return not self._wrapper_args['exclude_gradient']

    @contextlib.contextmanager
    def silent(self):
        save_stdout = sys.stdout
        sys.stdout = open(os.devnull, 'w')
        try:
            yield
        finally:
            sys.stdout = save_stdout


class MCOptimizer(Optimizer):
    """Scores ``nsamples`` uniformly random points and returns the best (gpflowopt/optim.py:131-172)."""

    def __init__(self, domain, nsamples, device_rng=False):
        """``device_rng=True`` (opt-in): for gpflowopt_b200 acquisitions the candidates are generated
        inside the scoring kernel from a counter-based stream (seeded from ``np.random``), every rank producing its own
        shard, so no N x D array is built, copied or broadcast; the winning row is regenerated on the host.  The points
        differ from ``np.random.rand``'s, which is why the default keeps the reference's host generation."""
        super(MCOptimizer, self).__init__(domain, exclude_gradient=True)
        self._nsamples = nsamples
        self._device_rng = device_rng
        self.set_initial(np.empty((0, self.domain.size)))

    @property
    def domain(self):
        return self._domain

    @domain.setter
    def domain(self, dom):
        self._domain = dom

    def _get_eval_points(self):
        return RandomDesign(self._nsamples, self.domain).generate()

    def _get_eval_rows(self, first_row, count, out=None):
        """Rows [first_row, first_row + count) of the candidate set, for the pipelined multi-GPU route (written into ``out``,
        e.g. a pinned staging buffer, when one is given).  Successive calls
        continue np.random's stream, so the chunks concatenate to exactly what one ``_get_eval_points()`` call draws: the same
        ``np.random.rand`` numbers through the same diagonal map (gpflowopt/design.py:55-70, 83-94).  The two ``in domain``
        asserts of ``Design.generate`` are not repeated per chunk: U[0, 1) mapped affinely onto the box is inside it by
        construction, and at 1e7 x 16 those two passes cost more host time than the draw itself."""
        design = RandomDesign(count, self.domain)
        scale, shift = (design.generative_domain >> self.domain).diagonal()
        if type(design).create_design is RandomDesign.create_design:
            from ._lib import numpy_stream_design_rows
            X = numpy_stream_design_rows(count, scale, shift, out)     # np.random's own stream, drawn natively (same bytes)
            if X is not None:
                return X
        X = design.create_design()
        np.multiply(X, scale, out=X)               # in place: the same products and sums, without two more N x D temporaries
        np.add(X, shift, out=X)
        if out is not None:
            out[...] = X
            return out
        return X

    def _optimize_device_rng(self, objective, acq):
        from ._lib import random_rows
        n, lower, upper = self._nsamples, self.domain.lower, self.domain.upper
        seed = dist.broadcast_int(np.random.randint(0, 2 ** 62))
        rank, ws = dist.world()
        lo, hi = dist.shard_bounds(n, rank, ws)
        if hi > lo:
            v, i = acq.argmax_random(seed, lo, hi - lo, lower, upper)
            i = i + lo if i >= 0 else -1
        else:
            v, i = 0.0, -1
        best, idx = dist.reduce_pairs(*dist.allgather_pair(v, i))
        objective.counter += n
        if idx < 0:
            return self._no_finite_score(n)
        x = random_rows(seed, idx, 1, lower, upper)
        return OptimizeResult(x=x, success=True, fun=np.array([[-best]]), nfev=n, message="OK", seed=seed, index=idx)

    def _no_finite_score(self, n):
        """Every candidate scored NaN (or there were none): nothing to return.  (np.argmin would hand back row 0 with a NaN
        score; here the stage reports failure so that callers do not evaluate the objective on an empty array.)"""
        return OptimizeResult(x=np.empty((0, self.domain.size)), success=False, fun=np.array([[np.nan]]), nfev=n,
                              message="No candidate received a finite acquisition score.")

    def _optimize(self, objective):
        acq = getattr(objective, 'acquisition', None)
        if acq is not None and self._device_rng and type(self)._get_eval_points is MCOptimizer._get_eval_points:
            return self._optimize_device_rng(objective, acq)
        if acq is not None:
            # fused route: argmin of -acq == argmax of acq with the same first-occurrence tie-break.  Under
            # torch.distributed only rank 0 produces the rows (the reference's single RNG stream); they are scattered in
            # pipelined rounds and every rank scores its chunks while rank 0 draws the next ones (dist.pipelined_argmax)
            custom = type(self)._get_eval_points not in (MCOptimizer._get_eval_points, CandidateOptimizer._get_eval_points)
            if custom:                                # a subclass that only overrides _get_eval_points: one draw, then slices
                held = self._get_eval_points() if dist.world()[0] == 0 else None
                produce = lambda first, count: held[first:first + count]
                n_total = held.shape[0] if held is not None else 0
            else:
                produce, n_total = self._get_eval_rows, self._nsamples
            engine = acq._get_engine() if hasattr(acq, '_get_engine') else None
            restore = []
            try:
                best, idx, x = dist.pipelined_argmax(
                    n_total, self.domain.size, produce, acq.argmax,
                    on_chunk_rows=(lambda rows: restore.append(engine.pin_variant_for(acq, rows))) if engine is not None else None)
            finally:
                if restore and engine is not None and engine.ctx is not None:
                    engine.ctx.set_variant(restore[0])
            n_total = self._nsamples if not custom else int(dist.broadcast_int(n_total))
            objective.counter += n_total
            if idx < 0:
                return self._no_finite_score(n_total)
            return OptimizeResult(x=x, success=True, fun=np.array([[-best]]), nfev=n_total, message="OK")
        points = self._get_eval_points()
        evaluations = objective(points)
        idx_best = np.argmin(evaluations, axis=0)
        return OptimizeResult(x=points[idx_best, :], success=True, fun=evaluations[idx_best, :],
                              nfev=points.shape[0], message="OK")

    def set_initial(self, initial):
        initial = np.atleast_2d(initial)
        if initial.size > 0:
            warnings.warn("Initial points set in {0} are ignored.".format(self.__class__.__name__), UserWarning)
            return
        super(MCOptimizer, self).set_initial(initial)


class CandidateOptimizer(MCOptimizer):
    """Scores a fixed candidate matrix (gpflowopt/optim.py:175-198)."""

    def __init__(self, domain, candidates):
        super(CandidateOptimizer, self).__init__(domain, candidates.shape[0])
        assert candidates in domain
        self.candidates = candidates

    def _get_eval_points(self):
        return self.candidates

    def _get_eval_rows(self, first_row, count, out=None):
        return self.candidates[first_row:first_row + count]

    @property
    def domain(self):
        return self._domain

    @domain.setter
    def domain(self, dom):
        t = self._domain >> dom
        self._domain = dom
        self.candidates = t.forward(self.candidates)


class SciPyOptimizer(Optimizer):
    """scipy.optimize.minimize wrapper (gpflowopt/optim.py:201-224); consumes the gradient kernel."""

    def __init__(self, domain, method='L-BFGS-B', tol=None, maxiter=1000):
        super(SciPyOptimizer, self).__init__(domain)
        self.config = dict(tol=tol, method=method, options=dict(disp=False, maxiter=maxiter))

    def _optimize(self, objective):
        want_jac = self.gradient_enabled()

        def objective1d(X):
            f, g = (arr.ravel() for arr in objective(X))
            if want_jac and g.size != np.size(X):
                # an empty gradient makes L-BFGS-B report convergence at the start point without taking a step
                raise ValueError("SciPyOptimizer needs the gradient: the objective returned %d of %d components"
                                 % (g.size, np.size(X)))
            return f, g
        config = dict(self.config)
        if not config['options'].get('disp'):          # recent SciPy warns about an explicit disp=False
            config['options'] = {k: v for k, v in config['options'].items() if k != 'disp'}
        return minimize(fun=objective1d, x0=self.get_initial().ravel(), jac=want_jac,
                        bounds=list(zip(self.domain.lower, self.domain.upper)), **config)


class _LockStep(object):
    """Rendezvous for k optimiser threads: every active thread posts its next point, the last one to arrive evaluates the
    whole batch with ONE objective call and hands each thread its row.  A run that finishes leaves the rendezvous."""

    def __init__(self, batch_fn, nruns):
        self._fn = batch_fn
        self._cv = threading.Condition()
        self._active = nruns
        self._pending = {}
        self._results = {}
        self._error = None
        self.batches = 0
        self.rows = 0

    def _flush(self):
        ids = sorted(self._pending)
        try:
            f, g = self._fn(np.vstack([self._pending[i] for i in ids]))
            for row, i in enumerate(ids):
                self._results[i] = (float(np.ravel(f[row])[0]), np.array(g[row], dtype=np.float64))
        except BaseException as exc:           # wake everybody up, every run re-raises
            self._error = exc
        self._pending.clear()
        self.batches += 1
        self.rows += len(ids)
        self._cv.notify_all()

    def evaluate(self, run, x):
        with self._cv:
            self._pending[run] = np.atleast_2d(x)
            if len(self._pending) == self._active:
                self._flush()
            while run not in self._results and self._error is None:
                self._cv.wait()
            if self._error is not None:
                raise self._error
            return self._results.pop(run)

    def leave(self, run):
        with self._cv:
            self._active -= 1
            if self._pending and len(self._pending) == self._active:
                self._flush()


class MultiStartOptimizer(Optimizer):
    """Monte-Carlo stage + batched multi-start polish (no counterpart class in the reference, whose
    StagedOptimizer([MCOptimizer, SciPyOptimizer]) polishes the single best MC point with N = 1 objective calls).

    The ``nstarts`` best of ``nsamples`` random points each seed an L-BFGS-B run; the runs advance in lock step, so every
    round of line-search evaluations is one ``evaluate_with_gradients`` call with N = (runs still active) instead of that
    many latency-bound N = 1 calls.  Each run sees exactly the numbers it would see alone, so its trajectory is the one
    ``SciPyOptimizer`` would produce from the same start; the result is the best run."""

    def __init__(self, domain, nsamples, nstarts=8, method='L-BFGS-B', tol=None, maxiter=1000, device_rng=False):
        super(MultiStartOptimizer, self).__init__(domain)
        self._nsamples = nsamples
        self._nstarts = nstarts
        self._device_rng = device_rng            # as MCOptimizer: candidates drawn inside the scoring kernel
        self.config = dict(tol=tol, method=method, options=dict(maxiter=maxiter))

    @staticmethod
    def _lowest(scores, k):
        """Indices of the k lowest scores, ascending, ties by index (what a stable argsort would put first)."""
        scores = np.ravel(scores)
        k = max(1, min(k, scores.size))
        if k < scores.size:
            kth = np.partition(scores, k - 1)[k - 1]
            cand = np.flatnonzero(scores <= kth)
        else:
            cand = np.arange(scores.size)
        return cand[np.argsort(scores[cand], kind='stable')][:k]

    def _starts(self, objective):
        acq = getattr(objective, 'acquisition', None)
        if acq is not None and self._device_rng:
            from ._lib import random_rows
            seed = int(np.random.randint(0, 2 ** 62))
            scores = -acq.evaluate_random(seed, 0, self._nsamples, self.domain.lower, self.domain.upper)
            objective.counter += self._nsamples
            order = self._lowest(scores, self._nstarts)
            points = np.vstack([random_rows(seed, int(i), 1, self.domain.lower, self.domain.upper) for i in order])
            return points, np.ravel(scores)[order]
        # the rows RandomDesign(n, domain).generate() would draw (same np.random stream, same affine map), produced natively
        # and without its two domain-membership passes over the N x D array (MCOptimizer._get_eval_rows)
        points = MCOptimizer(self.domain, self._nsamples)._get_eval_rows(0, self._nsamples)
        if acq is not None:                      # value-only scoring kernel; minimise -acq (bo.py:244-245)
            scores = -acq.evaluate(points)
            objective.counter += points.shape[0]
        else:
            scores = objective(points)[0]
        order = self._lowest(scores, self._nstarts)
        return points[order], np.ravel(scores)[order]

    def _optimize(self, objective):
        starts, start_scores = self._starts(objective)
        k = starts.shape[0]
        lock = _LockStep(objective, k)
        bounds = list(zip(self.domain.lower, self.domain.upper))
        results, errors = [None] * k, [None] * k

        def run(i):
            try:
                results[i] = minimize(fun=lambda x: lock.evaluate(i, x), x0=starts[i], jac=True, bounds=bounds, **self.config)
            except BaseException as exc:
                errors[i] = exc
            finally:
                lock.leave(i)

        threads = [threading.Thread(target=run, args=(i,), daemon=True) for i in range(k)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for exc in errors:
            if exc is not None:
                raise exc
        best = int(np.argmin([r.fun for r in results]))
        res = results[best]
        return OptimizeResult(x=res.x, fun=np.array([[res.fun]]), success=any(r.success for r in results),
                              message=res.message, nit=res.nit, nstarts=k, nbatches=lock.batches,
                              polish_rows=lock.rows, start_scores=start_scores, runs=results)


class StagedOptimizer(Optimizer):
    """Pipeline of optimisers; the best point of a stage seeds the next one (gpflowopt/optim.py:227-282)."""

    def __init__(self, optimizers):
        assert all(optimizers[0].domain == opt.domain for opt in optimizers)
        no_gradient = any(not opt.gradient_enabled() for opt in optimizers)
        super(StagedOptimizer, self).__init__(optimizers[0].domain, exclude_gradient=no_gradient)
        self.optimizers = optimizers
        del self._initial

    @property
    def domain(self):
        return self._domain

    @domain.setter
    def domain(self, domain):
        self._domain = domain
        for optimizer in self.optimizers:
            optimizer.domain = domain

    def _best_x(self, results):
        ok = [r for r in results if r.success]
        best_idx = int(np.argmin([np.ravel(r.fun)[0] for r in ok]))
        return ok[best_idx].x, ok[best_idx].fun

    def optimize(self, objectivefx):
        results = []
        result = None
        for current, following in zip(self.optimizers[:-1], self.optimizers[1:]):
            result = current.optimize(objectivefx)
            results.append(result)
            if not result.success:
                result.message += " StagedOptimizer interrupted after {0}.".format(current.__class__.__name__)
                break
            following.set_initial(self._best_x(results)[0])
        if result is None or result.success:
            result = self.optimizers[-1].optimize(objectivefx)
            results.append(result)
        result.nfev = sum(r.nfev for r in results)
        result.nstages = len(results)
        if any(r.success for r in results):
            result.x, result.fun = self._best_x(results)
        return result

    def get_initial(self):
        return self.optimizers[0].get_initial()

    def set_initial(self, initial):
        self.optimizers[0].set_initial(initial)
