"""Generate tests/golden/acq_golden.npz by running the REFERENCE's own ``build_acquisition`` methods
(ei.py:70-79, poi.py:63-67, pof.py:81-85, lcb.py:54-57, mes.py:96-102, hvpoi.py:98-140) on given predictive moments.

Run in the build container only (needs /root/reference):  python tests/golden/gen_acq_golden.py

What is executed is the reference's op-by-op composition (clamps, which moments enter where, gather indices of the cell
tables, reductions).  TensorFlow itself is absent, so its primitives are NumPy calls of the same meaning (``tf.maximum`` ->
``np.maximum``, ``tf.gather_nd`` -> integer indexing, ...) and ``tf.contrib.distributions.Normal`` evaluates with the
oracle's restatement of TF's ndtr / prob / log_ndtr (oracle/acq.py).  The fixture therefore pins the GPflowOpt-level
arithmetic of the oracle against the reference's own code; the TF-internal forms stay "recalled" (SURVEY Appendix B).
The Pareto cell tables come from the reference's own pareto.py, as in gen_pareto_golden.py.
"""
import importlib
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import acq as oacq            # noqa: E402  (Normal primitives only)


def numpy_tf(tf):
    def reduce(fn):
        return lambda x, axis=None, keep_dims=False, name=None: fn(x, axis=axis, keepdims=keep_dims)

    def gather_nd(params, indices):
        indices = np.asarray(indices)
        return np.asarray(params)[tuple(indices[:, k] for k in range(indices.shape[1]))]

    class Normal(object):
        def __init__(self, loc, scale):
            self.loc, self.scale = np.asarray(loc, dtype=float), np.asarray(scale, dtype=float)

        def cdf(self, x, name=None):
            return oacq.normal_cdf(np.asarray(x, dtype=float), self.loc, self.scale)

        def prob(self, x, name=None):
            return oacq.normal_prob(np.asarray(x, dtype=float), self.loc, self.scale)

        def log_cdf(self, x, name=None):
            return oacq.log_ndtr((np.asarray(x, dtype=float) - self.loc) / self.scale)

    tf.shape = lambda x: np.array(np.shape(x))
    tf.ones = lambda shape, dtype=None: np.ones([int(s) for s in shape])
    tf.concat = lambda values, axis: np.concatenate([np.asarray(v, dtype=float) for v in values], axis=axis)
    tf.maximum, tf.sqrt = np.maximum, np.sqrt
    tf.transpose = lambda x, perm=None: np.transpose(x, perm)
    tf.expand_dims = lambda x, axis: np.expand_dims(np.asarray(x), axis)
    tf.tile = lambda x, multiples: np.tile(x, [int(m) for m in multiples])
    tf.range = lambda n: np.arange(int(n))
    tf.stack = lambda values, axis=0: np.stack([np.asarray(v) for v in values], axis=axis)
    tf.reshape = lambda x, shape: np.reshape(x, [int(s) for s in shape])
    tf.gather_nd = gather_nd
    tf.reduce_sum, tf.reduce_prod, tf.reduce_all = reduce(np.sum), reduce(np.prod), reduce(np.all)
    tf.cast = lambda x, dtype=None: np.asarray(x).astype(float)
    tf.multiply = lambda a, b, name=None: a * b
    tf.add = lambda a, b, name=None: a + b
    tf.subtract = lambda a, b, name=None: a - b
    tf.constant = lambda v, dtype=None: np.asarray(v, dtype=float)
    tf.contrib = types.SimpleNamespace(distributions=types.SimpleNamespace(Normal=Normal))


def reference_cell_tables(cloud):
    """Front and cell bounds from the reference's own pareto.py (loaded as in gen_pareto_golden.py)."""
    spec = importlib.util.spec_from_file_location("gen_pareto_golden", os.path.join(HERE, "gen_pareto_golden.py"))
    gp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gp)
    mod = gp.load_reference_pareto()
    pareto = mod.Pareto(np.zeros((0, cloud.shape[1])))
    pareto.update(cloud)
    return np.array(pareto.front.value), np.array(pareto.bounds.lb.value), np.array(pareto.bounds.ub.value)


def main():
    rs = np.random.RandomState(20240614)
    R = 3
    cloud = rs.rand(30, R)
    front, lb, ub = reference_cell_tables(cloud)               # before the acquisition-side stand-ins replace the stubs
    spec = importlib.util.spec_from_file_location("gen_mes_golden", os.path.join(HERE, "gen_mes_golden.py"))
    stubs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(stubs)
    stubs._install_stubs()
    numpy_tf(sys.modules["tensorflow"])
    ref = {n: importlib.import_module("gpflowopt.acquisition." + n) for n in ("ei", "poi", "pof", "lcb", "mes", "hvpoi")}

    N = 40
    mean = rs.randn(N, 1) * 1.5
    var = np.abs(rs.randn(N, 1)) * 0.8
    var[:4] = [[1e-9], [0.0], [-1e-5], [1e-6]]                  # below / at the 1e-6 clamp, negative (LCB clamps at 0)
    out = {"mean": mean, "var": var}

    def probe(cls, models, **attrs):
        """The reference class without its constructor; its (property-backed) attributes become plain values."""
        members = {"models": property(lambda self: models)}
        members.update({k: property(lambda self, v=v: v) for k, v in attrs.items()})
        return object.__new__(type("Probe", (cls,), members))

    one = [types.SimpleNamespace(build_predict=lambda X: (mean, var))]
    fmin, threshold, sigma = np.array([-0.4]), 0.15, np.array(2.0)
    samples = np.array([-2.5, -2.0, -1.7, -3.1])
    out.update(fmin=fmin, threshold=np.array(threshold), sigma=sigma, samples=samples)
    out["ei"] = probe(ref["ei"].ExpectedImprovement, one, fmin=fmin).build_acquisition(None)
    out["poi"] = probe(ref["poi"].ProbabilityOfImprovement, one, fmin=fmin).build_acquisition(None)
    out["pof"] = probe(ref["pof"].ProbabilityOfFeasibility, one, threshold=threshold).build_acquisition(None)
    out["lcb"] = probe(ref["lcb"].LowerConfidenceBound, one, sigma=sigma).build_acquisition(None)
    with np.errstate(all="ignore"):                             # var <= 0 rows give nan / inf, as in the reference
        out["mes"] = probe(ref["mes"].MinValueEntropySearch, one, samples=samples, num_samples=samples.size).build_acquisition(None)

    # HVPoI: 3 objectives, front and cell tables from the reference's own Pareto class
    reference = np.max(front, axis=0, keepdims=True) + 0.1
    hmean, hvar = rs.rand(N, R) * 1.2 - 0.1, np.abs(rs.randn(N, R)) * 0.05
    hv_models = [types.SimpleNamespace(build_predict=lambda X, j=j: (hmean[:, [j]], hvar[:, [j]])) for j in range(R)]
    hv = probe(ref["hvpoi"].HVProbabilityOfImprovement, hv_models, data=(None, np.zeros((1, R))),
               pareto=types.SimpleNamespace(front=front, bounds=types.SimpleNamespace(lb=lb, ub=ub)), reference=reference)
    with np.errstate(invalid="ignore"):
        out["hvpoi"] = hv.build_acquisition(np.zeros((N, 2)))
    out.update(hv_mean=hmean, hv_var=hvar, hv_front=front, hv_lb=lb, hv_ub=ub, hv_reference=reference)
    np.savez(os.path.join(HERE, "acq_golden.npz"), **out)
    for k in ("ei", "poi", "pof", "lcb", "mes", "hvpoi"):
        print(k, np.asarray(out[k]).shape, np.asarray(out[k]).ravel()[:3])


if __name__ == "__main__":
    main()
