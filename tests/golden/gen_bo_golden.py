"""Generate tests/golden/bo_golden.npz by running the REFERENCE's own BayesianOptimizer._create_bo_result (bo.py:156-203).

Run in the build container only (needs /root/reference):  python tests/golden/gen_bo_golden.py

The method is NumPy only (feasibility mask, objective / constraint columns, arg-min or non-dominated set); it is called on an
instance created without its constructor whose ``acquisition`` is a plain record, behind the same GPflow / TensorFlow
stand-ins as gen_mes_golden.py.
"""
import importlib
import importlib.util
import os
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def cases():
    rs = np.random.RandomState(20240613)
    X = rs.rand(12, 3)
    Y1 = rs.randn(12, 1)
    Y2 = rs.rand(12, 2)
    Yc = rs.randn(12, 1)
    some = rs.rand(12) > 0.4
    return [
        dict(name="single", X=X, Y=Y1, feasible=np.ones(12, bool), obj=[0], con=[]),
        dict(name="pareto", X=X, Y=Y2, feasible=np.ones(12, bool), obj=[0, 1], con=[]),
        dict(name="constrained", X=X, Y=np.hstack((Y1, Yc)), feasible=some, obj=[0], con=[1]),
        dict(name="none_feasible", X=X, Y=np.hstack((Y1, Yc)), feasible=np.zeros(12, bool), obj=[0], con=[1]),
        dict(name="constraints_only", X=X, Y=Yc, feasible=some, obj=[], con=[0]),
    ]


def record(c):
    return types.SimpleNamespace(data=(c["X"], c["Y"]), feasible_data_index=lambda: c["feasible"],
                                 objective_indices=lambda: np.array(c["obj"], dtype=int),
                                 constraint_indices=lambda: np.array(c["con"], dtype=int))


def main():
    spec = importlib.util.spec_from_file_location("gen_mes_golden", os.path.join(HERE, "gen_mes_golden.py"))
    stubs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(stubs)
    stubs._install_stubs()
    import sys
    gpr = types.ModuleType("gpflow.gpr")
    gpr.GPR = type("GPR", (object,), {})
    sys.modules["gpflow.gpr"] = gpr
    sys.modules["gpflow"].gpr = gpr
    base = importlib.import_module("gpflowopt.acquisition.acquisition")     # the package shell has no __init__: expose the
    for name in ("Acquisition", "MCMCAcquistion"):                          # two names bo.py imports from the package
        setattr(sys.modules["gpflowopt.acquisition"], name, getattr(base, name))
    bo = importlib.import_module("gpflowopt.bo")
    out = {"names": np.array([c["name"] for c in cases()])}
    for c in cases():
        opt = object.__new__(bo.BayesianOptimizer)
        opt.acquisition = record(c)
        r = opt._create_bo_result(True, "OK")
        n = c["name"]
        for key in ("X", "Y", "feasible"):
            out["%s_%s" % (n, key)] = np.asarray(c[key])
        out["%s_obj" % n], out["%s_con" % n] = np.array(c["obj"], dtype=int), np.array(c["con"], dtype=int)
        out["%s_x" % n], out["%s_fun" % n], out["%s_constraints" % n] = np.atleast_2d(r.x), np.atleast_2d(r.fun), np.atleast_2d(r.constraints)
        out["%s_success" % n], out["%s_message" % n] = np.array(bool(r.success)), np.array(str(r.message))
        print(n, np.atleast_2d(r.x).shape, bool(r.success), r.message)
    np.savez(os.path.join(HERE, "bo_golden.npz"), **out)


if __name__ == "__main__":
    main()
