"""Generate tests/golden/pareto_golden.npz by running the REFERENCE's own gpflowopt/pareto.py.

Run in the build container only (needs /root/reference):  python tests/golden/gen_pareto_golden.py

The reference module imports GPflow 0.5 / TensorFlow 1 solely for container classes (Parameterized,
DataHolder) and the AutoFlow decorator of ``hypervolume``; the front / cell computations are NumPy.  We
therefore load the file as-is behind minimal stand-ins for those containers, so that the integer cell
tables (bounds.lb / bounds.ub) and fronts stored in the fixture are outputs of the reference code itself.
"""
import importlib.util
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

        def set_data(self, array):
            self._array = np.asarray(array)

    class Parameterized(object):
        def __setattr__(self, key, value):
            current = self.__dict__.get(key, None)
            if isinstance(current, DataHolder) and not isinstance(value, DataHolder):
                current.set_data(value)              # GPflow routes array assignment to DataHolder.set_data
                return
            object.__setattr__(self, key, value)

    def AutoFlow(*specs):
        def deco(fn):
            return fn
        return deco

    tf = types.ModuleType("tensorflow")
    tf.int32, tf.int64, tf.float64, tf.float32 = "int32", "int64", "float64", "float32"
    gpflow = types.ModuleType("gpflow")
    param = types.ModuleType("gpflow.param")
    param.Parameterized, param.DataHolder, param.AutoFlow = Parameterized, DataHolder, AutoFlow
    settings = types.SimpleNamespace(
        dtypes=types.SimpleNamespace(int_type=tf.int32, float_type=tf.float64),
        numerics=types.SimpleNamespace(jitter_level=1e-6))
    gpflow.param, gpflow.settings = param, settings
    sys.modules.update({"tensorflow": tf, "gpflow": gpflow, "gpflow.param": param})


def load_reference_pareto():
    _install_stubs()
    spec = importlib.util.spec_from_file_location("ref_pareto", os.path.join(REF, "gpflowopt", "pareto.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_cases():
    """Seeded objective clouds: R = 2, 3, 4; small and mid-sized fronts; one with duplicated rows."""
    rs = np.random.RandomState(20240607)
    cases = {}
    for name, (n, R) in {"r2_n30": (30, 2), "r2_n200": (200, 2), "r3_n25": (25, 3), "r3_n120": (120, 3),
                         "r4_n20": (20, 4)}.items():
        cases[name] = rs.rand(n, R)
    # points on a (noisy) sphere octant produce large fronts
    for name, (n, R) in {"r3_sphere60": (60, 3), "r2_sphere80": (80, 2)}.items():
        Z = np.abs(rs.randn(n, R))
        cases[name] = 1.0 - Z / np.linalg.norm(Z, axis=1, keepdims=True) * (1 + 0.01 * rs.rand(n, 1))
    dup = rs.rand(12, 3)
    cases["r3_dups"] = np.vstack((dup, dup[:4]))
    cases["test_pareto_py"] = np.array([[0.9575, 0.4218], [0.9649, 0.9157], [0.1576, 0.7922], [0.9706, 0.9595],
                                        [0.9572, 0.6557], [0.4854, 0.0357], [0.8003, 0.8491], [0.1419, 0.9340]])
    return cases


def main():
    ref = load_reference_pareto()
    out = {}
    for name, Y in make_cases().items():
        p = ref.Pareto(np.zeros((1, Y.shape[1])))
        p.update(Y)
        out[name + "/Y"] = Y
        out[name + "/front"] = p.front.value
        out[name + "/lb"] = p.bounds.lb.value
        out[name + "/ub"] = p.bounds.ub.value
        _, dom = ref.non_dominated_sort(Y)
        out[name + "/dominance"] = dom
        if Y.shape[1] == 2:                                   # also the generic strategy on 2-D data
            g = ref.Pareto(np.zeros((1, 2)))
            g.update(Y, generic_strategy=True)
            out[name + "/lb_generic"] = g.bounds.lb.value
            out[name + "/ub_generic"] = g.bounds.ub.value
        print(name, "front", p.front.value.shape, "cells", p.bounds.lb.value.shape)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pareto_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
