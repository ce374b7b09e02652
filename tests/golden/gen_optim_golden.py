"""Generate tests/golden/optim_golden.npz by running the REFERENCE's own optimisers (gpflowopt/optim.py: Optimizer.optimize,
MCOptimizer, CandidateOptimizer, StagedOptimizer; gpflowopt/objective.py:96-113).

Run in the build container only (needs /root/reference):  python tests/golden/gen_optim_golden.py

The optimisers are NumPy / SciPy code.  Stand-ins: the GPflow / TensorFlow containers of gen_mes_golden.py, ``tf.matmul`` /
``tf.transpose`` as NumPy calls (RandomDesign maps the unit-cube draws into the domain through them), and GPflow 0.5's
``model.ObjectiveWrapper`` base class (absent) restated from its documented behaviour: evaluate, remember the last point with
a finite gradient, replace non-finite gradients by zeros.  ``objectives()`` below is shared with the test.
"""
import importlib
import importlib.util
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def objectives():
    def bowl(X):
        X = np.atleast_2d(X)
        return np.sum((X - 0.2) ** 2, axis=1, keepdims=True), 2.0 * (X - 0.2)

    def bumpy(X):
        X = np.atleast_2d(X)
        return (np.sum(X ** 2 + 0.3 * np.sin(7.0 * X + 0.5), axis=1, keepdims=True), 2 * X + 2.1 * np.cos(7.0 * X + 0.5))

    class Interrupting(object):
        def __init__(self, after, f):
            self.after, self.f, self.seen = after, f, 0

        def __call__(self, X):
            if self.seen >= self.after:
                raise KeyboardInterrupt
            self.seen += np.atleast_2d(X).shape[0]
            return self.f(X)
    return bowl, bumpy, Interrupting


def run_all(optim, make_domain, out=None):
    """The same script for the reference (generator) and the product (test): returns {name: OptimizeResult}."""
    bowl, bumpy, Interrupting = objectives()
    dom = make_domain([-1.0, -1.0], [1.0, 2.0])
    res = {}
    np.random.seed(31)
    res["mc_bumpy"] = optim.MCOptimizer(dom, 200).optimize(bumpy)
    rs = np.random.RandomState(5)
    cand = np.vstack((-1 + rs.rand(40, 2) * [2.0, 3.0], [[0.2, 0.2]], -1 + rs.rand(10, 2) * [2.0, 3.0], [[0.2, 0.2]]))
    res["candidate_tie"] = optim.CandidateOptimizer(dom, cand).optimize(bowl)
    moved = optim.CandidateOptimizer(dom, cand)
    moved.domain = make_domain([0.0, 0.0], [4.0, 1.0])            # the setter rescales the candidates into the new box
    res["candidate_moved"] = moved.optimize(bowl)
    res["moved_candidates"] = moved.candidates
    # (SciPyOptimizer cannot be executed as is: it hands scipy.optimize.minimize a 2-D x0, which current SciPy rejects)
    np.random.seed(32)
    staged = optim.StagedOptimizer([optim.MCOptimizer(dom, 50), optim.MCOptimizer(dom, 30)])
    res["staged_mc"] = staged.optimize(bumpy)
    np.random.seed(33)
    res["staged_interrupted"] = optim.StagedOptimizer([optim.MCOptimizer(dom, 50), optim.MCOptimizer(dom, 30)]).optimize(
        Interrupting(50, bumpy))
    np.random.seed(34)
    res["mc_interrupted"] = optim.MCOptimizer(dom, 20).optimize(Interrupting(0, bumpy))
    return res


def flatten(res):
    out = {}
    for name, r in res.items():
        if isinstance(r, np.ndarray):
            out[name] = r
            continue
        x = np.atleast_2d(r.x)
        if x.dtype == object:                             # interrupted before any evaluation: "last good value" is None
            out[name + "_x_is_none"] = np.array(True)
        else:
            out[name + "_x"] = x
        out[name + "_success"] = np.array(bool(r.success))
        out[name + "_nfev"] = np.array(int(r.nfev))
        out[name + "_message"] = np.array(str(r.message))
        if "fun" in r:
            out[name + "_fun"] = np.atleast_2d(r.fun)
        if "nstages" in r:
            out[name + "_nstages"] = np.array(int(r.nstages))
    return out


def _old_numpy_stacking():
    """The reference targets NumPy 1.x, where np.vstack / np.hstack accepted a ``map`` object (domain.py:97, design.py:116)."""
    for name in ("vstack", "hstack"):
        orig = getattr(np, name)
        setattr(np, name, lambda tup, _orig=orig, **kw: _orig(tup if isinstance(tup, (list, tuple)) else list(tup), **kw))


def main():
    _old_numpy_stacking()
    spec = importlib.util.spec_from_file_location("gen_mes_golden", os.path.join(HERE, "gen_mes_golden.py"))
    stubs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(stubs)
    stubs._install_stubs()

    class GPflowObjectiveWrapper(object):                 # GPflow 0.5 model.ObjectiveWrapper, restated (absent here)
        def __init__(self, objective):
            self._objective = objective
            self._previous_x = None

        def __call__(self, x):
            f, g = self._objective(x)
            g_is_fin = np.isfinite(g)
            if np.all(g_is_fin):
                self._previous_x = x
                return f, g
            print("Warning: inf or nan in gradient: replacing with zeros")
            return f, np.where(g_is_fin, g, 0.)
    sys.modules["gpflow.model"].ObjectiveWrapper = GPflowObjectiveWrapper
    import types
    sys.modules["gpflow"].settings.verbosity = types.SimpleNamespace(optimisation_verb=False)
    optim = importlib.import_module("gpflowopt.optim")
    domain = importlib.import_module("gpflowopt.domain")

    def make_domain(lo, up):
        return np.sum([domain.ContinuousParameter("x%d" % i, l, u) for i, (l, u) in enumerate(zip(lo, up))])

    flat = flatten(run_all(optim, make_domain))
    np.savez(os.path.join(HERE, "optim_golden.npz"), **flat)
    for k in sorted(flat):
        if k.endswith(("_nfev", "_success", "_message", "_nstages")):
            print(k, flat[k])


if __name__ == "__main__":
    main()
