"""Generate tests/golden/boloop_golden.npz by running the REFERENCE's own BayesianOptimizer.optimize / _optimize loop
(bo.py:205-256 with Optimizer.optimize, optim.py:63-85) on recording stand-ins for the acquisition and the inner optimiser.

Run in the build container only (needs /root/reference):  python tests/golden/gen_boloop_golden.py

``script(...)`` is shared with the test, which runs it on gpflowopt_b200.bo.BayesianOptimizer.
"""
import importlib
import importlib.util
import os
import sys
import types

import numpy as np
from scipy.optimize import OptimizeResult

HERE = os.path.dirname(os.path.abspath(__file__))


class Acquisition(object):
    """Holds the evaluated data like a real acquisition and records what the loop does with it."""

    def __init__(self, X, Y, log):
        self.data, self.log, self._needs_setup = (X, Y), log, True
        self.models = [types.SimpleNamespace(wrapped="model-a"), types.SimpleNamespace(wrapped="model-b")]

    def set_data(self, X, Y):
        self.log.append("set_data %s %s last_x=%s" % (X.shape, Y.shape, np.round(X[-1], 6).tolist()))
        self.data, self._needs_setup = (X, Y), True
        return Y.shape[1]

    def evaluate_with_gradients(self, x):
        self.log.append("evaluate_with_gradients %s" % (np.shape(x),))
        self._needs_setup = False
        return np.sum(np.atleast_2d(x) ** 2, axis=1, keepdims=True), 2.0 * np.atleast_2d(x)

    def evaluate(self, x):
        self.log.append("evaluate %s" % (np.shape(x),))
        self._needs_setup = False
        return np.sum(np.atleast_2d(x) ** 2, axis=1, keepdims=True)

    def feasible_data_index(self):
        return np.ones(self.data[0].shape[0], dtype=bool)

    def objective_indices(self):
        return np.array([0])

    def constraint_indices(self):
        return np.array([], dtype=int)


class InnerOptimizer(object):
    """Proposes scripted points; evaluates the objective it is handed once, like a one-step optimiser."""

    def __init__(self, points, log):
        self.points, self.log, self.calls, self.domain = points, log, 0, None

    def gradient_enabled(self):
        return True

    def optimize(self, objective):
        x = np.atleast_2d(self.points[self.calls])
        self.calls += 1
        value = objective(x)
        f = value[0] if isinstance(value, tuple) else value
        self.log.append("inner optimize -> x=%s f=%s" % (np.round(x, 6).tolist(), np.round(np.ravel(f), 9).tolist()))
        return OptimizeResult(x=x, fun=f, success=True, message="OK", nfev=1)


def script(make_bo, make_domain):
    log = []
    dom = make_domain([-1.0, -1.0], [1.0, 1.0])
    X0 = np.array([[0.5, -0.5], [-0.25, 0.75]])
    fx = lambda X: np.sum((np.atleast_2d(X) - 0.1) ** 2, axis=1, keepdims=True)
    acq = Acquisition(X0, fx(X0), log)
    inner = InnerOptimizer(np.array([[0.3, 0.3], [0.0, 0.2], [0.1, 0.1]]), log)
    callback = lambda models: log.append("callback %s" % ([str(m) for m in models],))
    initial = np.array([[0.9, 0.9], [-0.9, 0.0]])
    bo = make_bo(dom, acq, inner, callback, initial)
    result = bo.optimize(fx, n_iter=3)
    out = {"log": np.array(log), "x": np.atleast_2d(result.x), "fun": np.atleast_2d(result.fun), "nfev": np.array(int(result.nfev)),
           "success": np.array(bool(result.success)), "message": np.array(str(result.message)),
           "initial_after": np.atleast_2d(bo.get_initial()), "data_X": acq.data[0], "data_Y": acq.data[1]}
    return out


def main():
    spec = importlib.util.spec_from_file_location("gen_optim_golden", os.path.join(HERE, "gen_optim_golden.py"))
    helper = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(helper)
    helper._old_numpy_stacking()
    spec = importlib.util.spec_from_file_location("gen_bo_golden", os.path.join(HERE, "gen_bo_golden.py"))
    gb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gb)
    spec = importlib.util.spec_from_file_location("gen_mes_golden", os.path.join(HERE, "gen_mes_golden.py"))
    stubs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(stubs)
    stubs._install_stubs()

    class GPflowObjectiveWrapper(object):                 # GPflow 0.5 model.ObjectiveWrapper, restated (absent here)
        def __init__(self, objective):
            self._objective, self._previous_x = objective, None

        def __call__(self, x):
            f, g = self._objective(x)
            if np.all(np.isfinite(g)):
                self._previous_x = x
                return f, g
            return f, np.where(np.isfinite(g), g, 0.)
    sys.modules["gpflow.model"].ObjectiveWrapper = GPflowObjectiveWrapper
    sys.modules["gpflow"].settings.verbosity = types.SimpleNamespace(optimisation_verb=False)
    gpr = types.ModuleType("gpflow.gpr")
    gpr.GPR = type("GPR", (object,), {})
    sys.modules["gpflow.gpr"] = gpr
    sys.modules["gpflow"].gpr = gpr
    base = importlib.import_module("gpflowopt.acquisition.acquisition")
    for name in ("Acquisition", "MCMCAcquistion"):
        setattr(sys.modules["gpflowopt.acquisition"], name, getattr(base, name))
    bo_mod = importlib.import_module("gpflowopt.bo")
    domain = importlib.import_module("gpflowopt.domain")

    def make_domain(lo, up):
        return np.sum([domain.ContinuousParameter("x%d" % i, l, u) for i, (l, u) in enumerate(zip(lo, up))])

    def make_bo(dom, acq, inner, callback, initial):
        bo = object.__new__(bo_mod.BayesianOptimizer)
        bo._domain, bo._initial, bo._wrapper_args = dom, initial, dict(exclude_gradient=True)
        bo.acquisition, bo.optimizer, bo._model_callback, bo.verbose, bo._scaling = acq, inner, callback, False, False
        return bo

    out = script(make_bo, make_domain)
    np.savez(os.path.join(HERE, "boloop_golden.npz"), **out)
    for line in out["log"]:
        print(line)
    print(out["x"], out["fun"], out["nfev"], out["message"], out["initial_after"].shape)


if __name__ == "__main__":
    main()
