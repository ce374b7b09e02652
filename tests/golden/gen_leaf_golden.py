"""Generate tests/golden/leaf_golden.npz by running the REFERENCE's own Acquisition bookkeeping on recording stand-in models:
``_optimize_models`` (restart protocol, acquisition.py:92-122), ``set_data`` (:145-169), ``enable_scaling`` (:127-143) and the
``setup_required`` decorator (:32-54).

Run in the build container only (needs /root/reference):  python tests/golden/gen_leaf_golden.py

``script(...)`` is shared with the test, which runs it on gpflowopt_b200's classes.
"""
import importlib
import importlib.util
import os
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class Model(object):
    """Records everything the acquisition does to a model.  ``runs``: scripted outcomes of successive optimize() calls, a
    number = objective value of a successful run, None = factorisation failure."""

    def __init__(self, name, width, runs, failure, log):
        object.__setattr__(self, "_s", dict(name=name, runs=list(runs), failure=failure, log=log, calls=0))
        object.__setattr__(self, "Y", types.SimpleNamespace(shape=(7, width)))
        object.__setattr__(self, "X", types.SimpleNamespace(shape=(7, 2)))

    def __setattr__(self, key, value):
        s = self._s
        if key in ("X", "Y"):
            s["log"].append("%s.%s <- %s first=%r" % (s["name"], key, np.shape(value), float(np.ravel(value)[0])))
        elif key == "input_transform":
            s["log"].append("%s.input_transform A=%s b=%s" % (s["name"], np.round(np.diag(np.asarray(value.A)), 12).tolist(),
                                                              np.round(np.asarray(value.b), 12).tolist()))
        else:
            s["log"].append("%s.%s = %r" % (s["name"], key, value))

    def randomize(self):
        self._s["log"].append("%s.randomize" % self._s["name"])

    def optimize(self):
        s = self._s
        outcome = s["runs"][s["calls"]]
        s["calls"] += 1
        s["log"].append("%s.optimize -> %r" % (s["name"], outcome))
        if outcome is None:
            raise s["failure"]("not positive definite")
        return types.SimpleNamespace(fun=outcome, x=np.array([float(s["calls"])]))

    def set_state(self, x):
        self._s["log"].append("%s.set_state %r" % (self._s["name"], float(np.ravel(x)[0])))


def script(make_leaf, failure, make_domain):
    """make_leaf(models, restarts) -> acquisition object without its constructor; returns {key: array}."""
    out = {}
    # restart protocol: best run wins (first on ties), failures are skipped, all failing raises
    log = []
    models = [Model("m0", 1, [3.0, 1.5, 1.5, 2.0], failure, log), Model("m1", 2, [None, 0.7, None, 0.9], failure, log)]
    leaf = make_leaf(models, 4)
    leaf._optimize_models()
    out["restarts_log"] = np.array(log)
    log = []
    leaf = make_leaf([Model("m0", 1, [None, None], failure, log)], 2)
    try:
        leaf._optimize_models()
        out["all_failed"] = np.array("no error")
    except RuntimeError as exc:
        out["all_failed"] = np.array(str(exc))
    log = []
    leaf = make_leaf([Model("m0", 1, [1.0], failure, log)], 0)
    leaf._optimize_models()
    out["zero_restarts_log"] = np.array(log + ["<end>"])
    # set_data: every model takes its own columns, extra columns are left to the caller
    log = []
    leaf = make_leaf([Model("m0", 1, [], failure, log), Model("m1", 2, [], failure, log)], 1)
    leaf._needs_setup = False
    Y = np.arange(7 * 5, dtype=float).reshape(7, 5)
    out["set_data_return"] = np.array(leaf.set_data(np.full((7, 2), 0.25), Y))
    out["set_data_flag"] = np.array(bool(leaf._needs_setup))
    out["set_data_log"] = np.array(log)
    return out


def main():
    spec = importlib.util.spec_from_file_location("gen_optim_golden", os.path.join(HERE, "gen_optim_golden.py"))
    helper = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(helper)
    helper._old_numpy_stacking()
    spec = importlib.util.spec_from_file_location("gen_mes_golden", os.path.join(HERE, "gen_mes_golden.py"))
    stubs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(stubs)
    stubs._install_stubs()
    import sys
    base = importlib.import_module("gpflowopt.acquisition.acquisition")
    domain = importlib.import_module("gpflowopt.domain")
    failure = sys.modules["tensorflow"].errors.InvalidArgumentError

    def make_leaf(models, restarts):
        probe = type("Probe", (base.Acquisition,), {"models": property(lambda self: models)})
        leaf = object.__new__(probe)
        object.__setattr__(leaf, "optimize_restarts", restarts)
        object.__setattr__(leaf, "_needs_setup", True)
        object.__setattr__(leaf, "_parent", None)
        return leaf

    out = script(make_leaf, failure, None)
    np.savez(os.path.join(HERE, "leaf_golden.npz"), **out)
    for k in sorted(out):
        print(k, out[k])


if __name__ == "__main__":
    main()
