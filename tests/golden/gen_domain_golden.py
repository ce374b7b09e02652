"""Generate tests/golden/domain_golden.npz by running the REFERENCE's own gpflowopt/domain.py and design.py.

Run in the build container only (needs /root/reference):  python tests/golden/gen_domain_golden.py

Only NumPy parts are executed: ``Domain.__rshift__`` (the coefficients A, b of the affine map between two boxes,
SURVEY 8a row B) and ``Domain.__contains__``.  GPflow / TensorFlow are imported by those
modules for container classes and for graph-building methods that are never called here, so they are replaced by minimal
stand-ins (as in gen_pareto_golden.py); the numbers stored in the fixture are outputs of the reference code itself.
"""
import importlib
import os
import sys
import types

import numpy as np

REF = os.environ.get("GPFO_REFERENCE", "/root/reference")


def _install_stubs():
    class DataHolder(object):
        def __init__(self, array, on_shape_change='raise'):
            self._array = np.asarray(array)

        @property
        def value(self):
            return self._array.copy()

        @property
        def shape(self):
            return self._array.shape

    class Parentable(object):
        def __init__(self):
            self._parent = None

    class Parameterized(Parentable):
        pass

    def AutoFlow(*specs):
        return lambda fn: fn

    tf = types.ModuleType("tensorflow")
    tf.int32, tf.int64, tf.float64, tf.float32 = "int32", "int64", "float64", "float32"
    gpflow = types.ModuleType("gpflow")
    param = types.ModuleType("gpflow.param")
    param.Parameterized, param.DataHolder, param.AutoFlow, param.Parentable = Parameterized, DataHolder, AutoFlow, Parentable
    gpflow.param = param
    gpflow.settings = types.SimpleNamespace(dtypes=types.SimpleNamespace(int_type=tf.int32, float_type=tf.float64),
                                            numerics=types.SimpleNamespace(jitter_level=1e-6))
    pkg = types.ModuleType("gpflowopt")                    # package shell: sub-modules load without gpflowopt/__init__.py
    pkg.__path__ = [os.path.join(REF, "gpflowopt")]
    sys.modules.update({"tensorflow": tf, "gpflow": gpflow, "gpflow.param": param, "gpflowopt": pkg})


def box_pairs():
    rs = np.random.RandomState(20240611)
    pairs = [(np.array([-5.0, 0.0]), np.array([10.0, 15.0]), np.zeros(2), np.ones(2)),        # Branin box -> unit cube
             (np.zeros(3), np.ones(3), np.array([-1.0, 2.0, 1e-3]), np.array([1.0, 2.5, 1e3])),
             (np.full(8, -5.0), np.full(8, 10.0), np.zeros(8), np.ones(8))]
    for d in (1, 4, 16):
        lo1, w1 = rs.randn(d) * 3, rs.rand(d) * 10 + 0.1
        lo2, w2 = rs.randn(d), rs.rand(d) + 0.05
        pairs.append((lo1, lo1 + w1, lo2, lo2 + w2))
    return pairs


def main():
    _install_stubs()
    domain = importlib.import_module("gpflowopt.domain")

    def make(lo, up, tag):
        return np.sum([domain.ContinuousParameter("%s%d" % (tag, i), l, u) for i, (l, u) in enumerate(zip(lo, up))])

    out = {}
    pairs = box_pairs()
    out["n_pairs"] = np.array(len(pairs))
    for k, (lo1, up1, lo2, up2) in enumerate(pairs):
        t = make(lo1, up1, "a") >> make(lo2, up2, "b")
        out["pair%d_bounds" % k] = np.vstack((lo1, up1, lo2, up2))
        out["pair%d_A" % k] = t.A.value
        out["pair%d_b" % k] = t.b.value
    # membership: rows around the bounds of the Branin box (inside, on, within / beyond np.isclose of the bounds, wrong width)
    lo, up = np.array([-5.0, 0.0]), np.array([10.0, 15.0])
    dom = make(lo, up, "x")
    probes = np.array([[0.0, 1.0], [-5.0, 0.0], [10.0, 15.0], [-5.0 - 1e-9, 0.0], [10.0 + 1e-7, 15.0], [10.0 + 1e-3, 15.0],
                       [-5.1, 3.0], [2.0, 15.0 + 1e-12], [2.0, -1e-9], [2.0, -1e-6]])
    out["contains_bounds"] = np.vstack((lo, up))
    out["contains_probes"] = probes
    out["contains_rows"] = np.array([bool(p[None, :] in dom) for p in probes])
    out["contains_all"] = np.array(bool(probes[:3] in dom))
    out["contains_wrong_width"] = np.array(bool(np.zeros((2, 3)) in dom))
    # (FactorialDesign.create_design cannot be executed as is: np.vstack(map(...)) is rejected by current NumPy)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "domain_golden.npz")
    np.savez(path, **out)
    print("wrote", path, {k: np.asarray(v).shape for k, v in out.items() if k.startswith(("pair0", "contains_rows"))})


if __name__ == "__main__":
    main()
