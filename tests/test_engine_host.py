"""Host logic of the engine without a GPU: a recording stand-in for the ctypes Context shows WHAT the Python layer sends
through the C-ABI -- which models are (re-)uploaded when, slot assignment, program caching, setup order in trees
(constraints before objectives), set_data column slicing.  The numerical work itself is covered by the -m gpu tests."""
import numpy as np
import pytest

import gpflowopt_b200 as gpfo
from gpflowopt_b200 import _lib, engine as engine_mod
from gpflowopt_b200.acquisition import (AcquisitionProduct, AcquisitionSum, ExpectedImprovement, LowerConfidenceBound,
                                        ProbabilityOfFeasibility)
from gpflowopt_b200.domain import ContinuousParameter


class FakeContext(object):
    """Records calls; predictions are a deterministic function of the uploaded Y so that setup logic has numbers."""

    def __init__(self, device=0):
        self.calls = []
        self.gps = {}

    def close(self):
        pass

    def gp_set(self, gp_id, **kw):
        self.calls.append(("gp_set", gp_id, kw["slot0"], kw["X"].shape, float(np.sum(kw["Y"]))))
        self.gps[gp_id] = kw

    def gp_count_set(self, n):
        self.calls.append(("gp_count_set", n))

    def program_set(self, ops, params):
        self.calls.append(("program_set", [tuple(int(v) for v in o) for o in ops], [float(p) for p in params]))

    def cells_set(self, pf_ext, lb, ub):
        self.calls.append(("cells_set", pf_ext.shape, lb.shape))

    def predict(self, gp_id, X, R):
        y = self.gps[gp_id]["Y"]
        mean = np.tile(np.linspace(y.min(), y.max(), X.shape[0]).reshape(-1, 1), (1, R))
        return mean, np.full((X.shape[0], R), 0.1)

    def eval(self, X, want_values=True, want_grads=False):
        self.calls.append(("eval", X.shape, want_values, want_grads))
        vals = np.linspace(0.0, 1.0, X.shape[0]).reshape(-1, 1) if want_values else None
        grads = np.zeros(X.shape) if want_grads else None
        return vals, grads, 1.0, X.shape[0] - 1


@pytest.fixture
def fake(monkeypatch):
    monkeypatch.setattr(_lib, "Context", FakeContext)
    return FakeContext


def _data(n=12, seed=0):
    rs = np.random.RandomState(seed)
    X = rs.rand(n, 2) * 2 - 1
    return X, np.sum(X ** 2, axis=1, keepdims=True), X[:, [0]] - 0.2


def _gpr(X, Y):
    return gpfo.GPR(X, Y, gpfo.RBF(2, ARD=True, lengthscales=[0.5, 0.5]), noise_variance=1e-3)


def _names(ctx):
    return [c[0] for c in ctx.calls]


def test_uploads_only_when_state_changes(fake):
    X, Yo, _ = _data()
    ei = ExpectedImprovement(_gpr(X, Yo))
    ctx = ei._get_engine().ctx
    assert _names(ctx).count("gp_set") == 1                       # __init__ -> _setup -> predict_f factorises once
    Xc = np.random.RandomState(1).rand(7, 2)
    ei.evaluate(Xc)
    ei.evaluate(Xc)
    assert _names(ctx).count("gp_set") == 1 and _names(ctx).count("program_set") == 1 and _names(ctx).count("eval") == 2
    ei.models[0].wrapped.kern.lengthscales = [0.3, 0.3]            # hyper-parameter change -> one re-factorisation
    ei.evaluate(Xc)
    assert _names(ctx).count("gp_set") == 2
    Xn, Yn, _ = _data(15, seed=3)
    assert ei.set_data(Xn, Yn) == 1 and ei._needs_setup            # new data -> lazy setup re-armed
    ei.evaluate(Xc)
    assert not ei._needs_setup and _names(ctx).count("gp_set") == 3
    assert ctx.calls[-2][0] == "program_set"                        # fmin changed -> new program parameters
    v, g = ei.evaluate_with_gradients(Xc)
    assert v.shape == (7, 1) and g.shape == (7, 2) and ctx.calls[-1] == ("eval", (7, 2), True, True)


def test_tree_program_slots_and_setup_order(fake):
    X, Yo, Yc = _data()
    ei, pof, lcb = ExpectedImprovement(_gpr(X, Yo)), ProbabilityOfFeasibility(_gpr(X, Yc), threshold=0.1), \
        LowerConfidenceBound(_gpr(X, Yo), 1.5)
    tree = ei * (pof + lcb)
    assert isinstance(tree, AcquisitionProduct) and isinstance(tree.operands[1], AcquisitionSum)
    np.testing.assert_array_equal(tree.constraint_indices(), [1])
    np.testing.assert_array_equal(tree.objective_indices(), [0, 2])
    order = []
    for node, name in ((ei, "ei"), (pof, "pof")):
        orig = node._setup
        node._setup = (lambda o=orig, n=name: (order.append(n), o())[1])
    tree.evaluate(np.zeros((3, 2)))
    assert order[:2] == ["pof", "ei"]                               # constraints first (acquisition.py:334-347)
    ctx = tree._get_engine().ctx
    prog = [c for c in ctx.calls if c[0] == "program_set"][-1]
    assert prog[1] == [(_lib.OP_EI, 0, 0, 0), (_lib.OP_POF, 1, 0, 1), (_lib.OP_LCB, 2, 0, 2), (_lib.OP_SUM, 2, 0, 0),
                       (_lib.OP_PROD, 2, 0, 0)]
    assert prog[2][1:] == [0.1, 1.5]
    slots = sorted({(c[1], c[2]) for c in ctx.calls if c[0] == "gp_set"})
    assert slots == [(0, 0), (1, 1), (2, 2)]
    # set_data slices Y column-wise over the operands (acquisition.py:145-169, 328-332)
    Xn, Yon, Ycn = _data(14, seed=5)
    Y = np.hstack((Yon, Ycn, 2 * Yon))
    assert tree.set_data(Xn, Y) == 3
    np.testing.assert_allclose(pof.data[1], Ycn)
    np.testing.assert_allclose(lcb.data[1], 2 * Yon)
    # a sub-acquisition evaluated on its own (as PoF.feasible_data_index does) gets its own program on the same engine
    pof.evaluate(np.zeros((2, 2)))
    assert [c for c in ctx.calls if c[0] == "program_set"][-1][1] == [(_lib.OP_POF, 1, 0, 0)]


def test_enable_scaling_sends_transforms(fake):
    X, Yo, _ = _data()
    domain = ContinuousParameter("a", -1, 1) + ContinuousParameter("b", -1, 1)
    ei = ExpectedImprovement(_gpr(X, 3.0 * Yo + 1.0))
    ei.enable_scaling(domain)
    ei.evaluate(np.zeros((2, 2)))
    kw = ei._get_engine().ctx.gps[0]
    np.testing.assert_allclose(kw["in_scale"], [0.5, 0.5])
    np.testing.assert_allclose(kw["in_shift"], [0.5, 0.5])
    Y = 3.0 * Yo + 1.0
    np.testing.assert_allclose(kw["out_scale"], [1.0 / Y.std()])
    np.testing.assert_allclose(kw["out_shift"], [-Y.mean() / Y.std()])
    assert kw["X"].min() >= 0.0 and kw["X"].max() <= 1.0 and abs(kw["Y"].mean()) < 1e-12   # scaled data is what is uploaded
    np.testing.assert_allclose(ei.data[0], X)                                              # ... while .data stays unscaled
    np.testing.assert_allclose(ei.data[1], Y)


def test_unsupported_models_are_refused():
    with pytest.raises(AssertionError):
        ExpectedImprovement(object())
    from gpflowopt_b200.transforms import LinearTransform
    with pytest.raises(ValueError):
        LinearTransform(np.array([[1.0, 0.3], [0.0, 1.0]]), [0.0, 0.0]).diagonal()


# ---- round-2 regressions: gradients reach the SciPy stage of a staged optimiser, MCMC copies stay hidden, --------
# ---- k-ary aggregations keep the device stack shallow, all-NaN scores fail the Monte-Carlo stage -----------------------------
def _domain2():
    return ContinuousParameter("a", -1, 1) + ContinuousParameter("b", -1, 1)


def test_staged_optimizer_polish_stage_receives_gradients(fake):
    """BayesianOptimizer(optimizer=StagedOptimizer([MCOptimizer, SciPyOptimizer])): the staged optimiser as a whole is
    'gradient-free' (one stage is), but its SciPy stage must still be fed evaluate_with_gradients -- as the reference's
    inverse_acquisition always does (gpflowopt/bo.py:244-245).  An empty gradient makes L-BFGS-B stop at its start."""
    from gpflowopt_b200 import BayesianOptimizer
    from gpflowopt_b200.optim import MCOptimizer, SciPyOptimizer, StagedOptimizer
    X, Yo, _ = _data()
    dom = _domain2()
    acq = ExpectedImprovement(_gpr(X, Yo))
    acq.optimize_restarts = 0
    bo = BayesianOptimizer(dom, acq, optimizer=StagedOptimizer([MCOptimizer(dom, 50), SciPyOptimizer(dom)]), callback=None)
    np.random.seed(0)
    bo.optimize(lambda x: np.sum(np.atleast_2d(x) ** 2, axis=1, keepdims=True), n_iter=1)
    evals = [c for c in acq._get_engine().ctx.calls if c[0] == "eval"]
    assert ("eval", (50, 2), False, False) in evals                     # MC stage: fused argmax, no values, no gradients
    polish = [c for c in evals if c[1] == (1, 2)]
    assert polish and all(c[3] for c in polish)                         # every N = 1 polish call asks for gradients


def test_scipy_optimizer_refuses_an_empty_gradient():
    from gpflowopt_b200.optim import SciPyOptimizer
    opt = SciPyOptimizer(_domain2())
    with pytest.raises(ValueError, match="needs the gradient"):
        opt.optimize(lambda x: (np.sum(x ** 2, axis=1, keepdims=True), np.zeros((x.shape[0], 0))))


def test_mcmc_acquisition_hides_its_copies_and_runs_in_bo(fake):
    """MCMCAcquistion(acq, k): ``models`` / ``data`` show the first copy only (gpflowopt/acquisition/acquisition.py:434-437) so
    BayesianOptimizer(hyper_draws=k) sees ONE output column; the device still receives k GP states, and the program is
    folded pairwise (stack depth 2), so k = 10 draws fit the 8-deep device stack."""
    from gpflowopt_b200 import BayesianOptimizer
    from gpflowopt_b200.acquisition import MCMCAcquistion
    from gpflowopt_b200.optim import MCOptimizer
    X, Yo, _ = _data()
    dom = _domain2()
    k = 10
    bo = BayesianOptimizer(dom, LowerConfidenceBound(_gpr(X, Yo), 2.0), optimizer=MCOptimizer(dom, 40), hyper_draws=k,
                           callback=None)
    mc = bo.acquisition
    assert isinstance(mc, MCMCAcquistion) and len(mc.operands) == k
    assert len(mc.models) == 1 and len(mc._all_models()) == k
    assert mc.data[1].shape == (12, 1)
    np.testing.assert_array_equal(mc.objective_indices(), [0])
    np.random.seed(0)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = bo.optimize(lambda x: np.sum(np.atleast_2d(x) ** 2, axis=1, keepdims=True), n_iter=2)
    assert r.success and mc.data[0].shape[0] == 14 and all(o.data[0].shape[0] == 14 for o in mc.operands)
    ctx = mc._get_engine().ctx
    assert sorted({c[1] for c in ctx.calls if c[0] == "gp_set"}) == list(range(k))
    prog = [c for c in ctx.calls if c[0] == "program_set"][-1][1]
    depth = peak = 0
    for op in prog:
        if op[0] in (_lib.OP_SUM, _lib.OP_PROD):
            depth -= op[1] - 1
        elif op[0] != _lib.OP_SCALE:
            depth += 1
        peak = max(peak, depth)
    assert depth == 1 and peak == 2 and prog[-1][0] == _lib.OP_SCALE
    assert sum(1 for op in prog if op[0] == _lib.OP_LCB) == k


def test_many_operand_sum_is_folded_left_to_right(fake):
    X, Yo, _ = _data()
    leaves = [LowerConfidenceBound(_gpr(X, Yo + i), 1.0 + i) for i in range(4)]
    tree = leaves[0] + leaves[1] + leaves[2] + leaves[3]
    assert isinstance(tree, AcquisitionSum) and len(tree.operands) == 4
    tree.evaluate(np.zeros((2, 2)))
    prog = [c for c in tree._get_engine().ctx.calls if c[0] == "program_set"][-1][1]
    kinds = [op[0] for op in prog]
    assert kinds == [_lib.OP_LCB, _lib.OP_LCB, _lib.OP_SUM, _lib.OP_LCB, _lib.OP_SUM, _lib.OP_LCB, _lib.OP_SUM]
    assert all(op[1] == 2 for op in prog if op[0] == _lib.OP_SUM)


def test_mc_optimizer_reports_failure_when_no_score_is_finite(fake, monkeypatch):
    from gpflowopt_b200.optim import MCOptimizer
    X, Yo, _ = _data()
    dom = _domain2()
    acq = ExpectedImprovement(_gpr(X, Yo))
    monkeypatch.setattr(FakeContext, "eval", lambda self, X, want_values=True, want_grads=False: (None, None, 0.0, -1))

    def target(x):
        return -acq.evaluate(x), np.zeros((x.shape[0], 0))
    target.acquisition = acq
    r = MCOptimizer(dom, 30).optimize(target)
    assert not r.success and r.x.shape == (0, 2) and "finite" in r.message
