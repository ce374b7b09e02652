"""Generate tests/golden/mes_golden.npz by running the REFERENCE's own MinValueEntropySearch._setup (mes.py:63-94).

Run in the build container only (needs /root/reference):  python tests/golden/gen_mes_golden.py

``_setup`` is host-side NumPy / SciPy: a random grid over the domain, the model's predictive moments on data + grid, a
bracketed bisection for three quantiles of the distribution of the minimum, a Gumbel fit and ``num_samples`` draws.  The
model is replaced by an analytic stand-in (``surrogate_moments`` below, also used by the test), GPflow / TensorFlow by
minimal stand-ins whose only computing parts are ``tf.matmul`` / ``tf.transpose`` as NumPy calls (the reference maps the
unit-cube grid into the domain through them).  The stored samples are therefore outputs of the reference's own code.
"""
import importlib
import os
import sys
import types

import numpy as np

REF = os.environ.get("GPFO_REFERENCE", "/root/reference")


def surrogate_moments(X):
    """Smooth stand-in for predict_f: mean N x 1, variance N x 1."""
    X = np.atleast_2d(X)
    mean = np.sum((X - 0.3) ** 2, axis=1, keepdims=True) - 0.5 * np.cos(3.0 * X[:, [0]])
    var = 0.05 + 0.1 * np.sum(np.sin(2.0 * X) ** 2, axis=1, keepdims=True)
    return mean, var


def _install_stubs():
    class DataHolder(np.ndarray):
        def __new__(cls, array, on_shape_change='raise'):
            return np.asarray(array, dtype=float).view(cls)

        @property
        def value(self):
            return np.array(self)

    class Parentable(object):
        def __init__(self):
            self._parent = None

        @property
        def name(self):
            return type(self).__name__

        @property
        def highest_parent(self):
            return self

    class Parameterized(Parentable):
        pass

    def AutoFlow(*specs):
        return lambda fn: fn

    class ParamList(list):
        sorted_params = property(lambda self: list(self))

    tf = types.ModuleType("tensorflow")
    tf.int32, tf.int64, tf.float64, tf.float32 = "int32", "int64", "float64", "float32"
    tf.matmul, tf.transpose = np.matmul, np.transpose
    tf.reduce_sum, tf.reduce_prod = "reduce_sum", "reduce_prod"       # only stored as the aggregation's tag
    tf.errors = types.SimpleNamespace(InvalidArgumentError=RuntimeError)
    gpflow = types.ModuleType("gpflow")
    param = types.ModuleType("gpflow.param")
    for name, obj in dict(Parameterized=Parameterized, DataHolder=DataHolder, AutoFlow=AutoFlow, Parentable=Parentable,
                          ParamList=ParamList, Param=DataHolder).items():
        setattr(param, name, obj)
    model = types.ModuleType("gpflow.model")
    model.Model = type("Model", (Parameterized,), {})
    model.ObjectiveWrapper = type("ObjectiveWrapper", (object,), {})
    gpflow.param, gpflow.model = param, model
    gpflow.settings = types.SimpleNamespace(dtypes=types.SimpleNamespace(int_type=tf.int32, float_type=tf.float64),
                                            numerics=types.SimpleNamespace(jitter_level=1e-6))
    pkg = types.ModuleType("gpflowopt")
    pkg.__path__ = [os.path.join(REF, "gpflowopt")]
    acq_pkg = types.ModuleType("gpflowopt.acquisition")
    acq_pkg.__path__ = [os.path.join(REF, "gpflowopt", "acquisition")]
    sys.modules.update({"tensorflow": tf, "gpflow": gpflow, "gpflow.param": param, "gpflow.model": model,
                        "gpflowopt": pkg, "gpflowopt.acquisition": acq_pkg})
    return DataHolder


def cases():
    rs = np.random.RandomState(20240612)
    out = []
    for seed, d, n, lo, up in ((11, 2, 16, -1.0, 1.0), (12, 3, 25, 0.0, 1.0), (13, 2, 9, -2.0, 3.0)):
        X = lo + (up - lo) * rs.rand(n, d)
        out.append(dict(seed=seed, X=X, lower=np.full(d, lo), upper=np.full(d, up), gridsize=500, num_samples=10))
    return out


def main():
    DataHolder = _install_stubs()
    domain = importlib.import_module("gpflowopt.domain")
    mes = importlib.import_module("gpflowopt.acquisition.mes")

    class Probe(mes.MinValueEntropySearch):              # the reference's class with plain attributes behind its properties
        models = property(lambda self: self._probe_models)
        data = property(lambda self: self._probe_data)

    class Samples(object):
        def set_data(self, values):
            self.values = np.array(values)

    out = {"n_cases": np.array(len(cases()))}
    for k, c in enumerate(cases()):
        probe = object.__new__(Probe)
        probe._probe_models = [types.SimpleNamespace(predict_f=surrogate_moments)]
        probe._probe_data = (c["X"], surrogate_moments(c["X"])[0])
        probe.gridsize, probe.num_samples = c["gridsize"], c["num_samples"]
        probe._domain = np.sum([domain.ContinuousParameter("x%d" % i, l, u) for i, (l, u) in enumerate(zip(c["lower"], c["upper"]))])
        object.__setattr__(probe, "samples", Samples())
        np.random.seed(c["seed"])
        probe._setup()
        for key in ("seed", "X", "lower", "upper", "gridsize", "num_samples"):
            out["case%d_%s" % (k, key)] = np.asarray(c[key])
        out["case%d_samples" % k] = probe.samples.values
        print("case", k, "samples", np.round(probe.samples.values, 4))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mes_golden.npz")
    np.savez(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
