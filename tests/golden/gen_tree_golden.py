"""Generate tests/golden/tree_golden.npz by running the REFERENCE's own AcquisitionAggregation bookkeeping
(acquisition.py:328-358: set_data column offsets, _setup order, constraint_indices, feasible_data_index) on recording
stand-in operands.

Run in the build container only (needs /root/reference):  python tests/golden/gen_tree_golden.py

``script(cls)`` is shared with the test, which runs it on gpflowopt_b200's class.  Note what the fixture pins for three or more
operands: the reference shifts operand k's constraint columns by the width of operand k-1, not by the cumulative width
(acquisition.py:349-355) -- a drop-in has to reproduce that.
"""
import importlib
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class Operand(object):
    """Stands in for an acquisition: ``width`` output columns, local constraint columns, a feasibility mask; records calls."""

    def __init__(self, name, width, constraints, feasible, log):
        self.name, self.width, self._con, self._feasible, self.log = name, width, np.array(constraints, dtype=int), feasible, log
        self.data = (None, np.zeros((len(feasible), width)))

    def set_data(self, X, Y):
        self.log.append(("set_data", self.name, Y.shape[1], float(Y[0, 0])))
        return self.width

    def constraint_indices(self):
        return self._con

    def feasible_data_index(self):
        return self._feasible

    def _setup_constraints(self):
        self.log.append(("constraints", self.name))

    def _setup_objectives(self):
        self.log.append(("objectives", self.name))


def script(make_aggregation):
    rs = np.random.RandomState(20240616)
    n = 9
    out = {}
    for case, spec in enumerate(([(1, [], 0.0), (1, [0], 0.4)],
                                 [(2, [], 0.0), (1, [0], 0.3), (3, [1, 2], 0.5)],
                                 [(1, [0], 0.5), (2, [1], 0.2), (1, [], 0.0), (2, [0, 1], 0.6)])):
        log = []
        ops = [Operand("op%d" % k, w, con, rs.rand(n) >= p, log) for k, (w, con, p) in enumerate(spec)]
        agg = make_aggregation(ops)
        total = sum(w for w, _, _ in spec)
        Y = np.arange(n * total, dtype=float).reshape(n, total)
        out["case%d_set_data_return" % case] = np.array(agg.set_data(np.zeros((n, 2)), Y))
        agg._setup()
        out["case%d_log" % case] = np.array([" ".join(str(v) for v in entry) for entry in log])
        out["case%d_constraint_indices" % case] = np.asarray(agg.constraint_indices())
        out["case%d_feasible" % case] = np.asarray(agg.feasible_data_index())
    out["n_cases"] = np.array(3)
    return out


def operator_script(make_leaf):
    """Shapes of the trees the + and * operators build (flattening of equal aggregations, acquisition.py:269-293, 377-393)."""
    a, b, c, d = (make_leaf(tag) for tag in "abcd")

    def show(node):
        if hasattr(node, "operands"):
            return "%s(%s)" % (type(node).__name__, ", ".join(show(x) for x in node.operands))
        return node.tag
    exprs = ("a+b", "(a+b)+c", "a+(b+c)", "a*b*c", "(a+b)*c", "a+(b*c)", "(a*b)+(c*d)", "(a+b)*(c+d)", "a*(b*c)", "(a+b)+(c+d)")
    return np.array(["%s -> %s" % (e, show(eval(e))) for e in exprs])


def main():
    spec = importlib.util.spec_from_file_location("gen_optim_golden", os.path.join(HERE, "gen_optim_golden.py"))
    helper = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(helper)
    helper._old_numpy_stacking()
    spec = importlib.util.spec_from_file_location("gen_mes_golden", os.path.join(HERE, "gen_mes_golden.py"))
    stubs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(stubs)
    stubs._install_stubs()
    base = importlib.import_module("gpflowopt.acquisition.acquisition")

    def make_aggregation(ops):
        agg = object.__new__(type("Probe", (base.AcquisitionAggregation,), {}))
        object.__setattr__(agg, "operands", list(ops))
        return agg

    out = script(make_aggregation)

    def make_leaf(tag):
        leaf = object.__new__(type("Leaf", (base.Acquisition,), {}))
        object.__setattr__(leaf, "tag", tag)
        object.__setattr__(leaf, "_parent", None)
        return leaf
    out["operators"] = operator_script(make_leaf)
    print(out["operators"])
    np.savez(os.path.join(HERE, "tree_golden.npz"), **out)
    for k in sorted(out):
        if "constraint_indices" in k or "return" in k:
            print(k, out[k])
    print(out["case1_log"])


if __name__ == "__main__":
    main()
