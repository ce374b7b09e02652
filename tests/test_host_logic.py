"""Host-side mirror of the reference interface (no GPU): domains, transforms, designs, optimisers on plain Python
objectives, program construction, sharding logic.  Modelled on testing/unit/test_optimizers.py, test_domain.py,
test_transforms.py and test_acquisition.py of the reference."""
import warnings

import numpy as np
import pytest

import gpflowopt_b200 as gpfo
from gpflowopt_b200 import dist, _lib
from gpflowopt_b200.design import EmptyDesign, FactorialDesign, RandomDesign
from gpflowopt_b200.domain import ContinuousParameter, UnitCube
from gpflowopt_b200.engine import ProgramBuilder
from gpflowopt_b200.optim import CandidateOptimizer, MCOptimizer, MultiStartOptimizer, SciPyOptimizer, StagedOptimizer
from gpflowopt_b200.transforms import LinearTransform


def parabola2d(X):
    return np.atleast_2d(np.sum(X ** 2, axis=1)).T, 2 * X


class KeyboardRaiser(object):
    def __init__(self, iters_to_raise, f):
        self.iters_to_raise, self.f, self.count = iters_to_raise, f, 0

    def __call__(self, X):
        if self.count >= self.iters_to_raise:
            raise KeyboardInterrupt
        val = self.f(X)
        self.count += X.shape[0]
        return val


def _domain():
    return ContinuousParameter("x1", -1.0, 1.0) + ContinuousParameter("x2", -1.0, 1.0)


def test_domain_transform_and_contains():
    d = _domain()
    t = d >> UnitCube(2)
    np.testing.assert_allclose(np.diag(t.A), [0.5, 0.5])
    np.testing.assert_allclose(t.b, [0.5, 0.5])
    assert np.array([[1.0, -1.0]]) in d and np.array([[1.1, 0.0]]) not in d and np.zeros((1, 3)) not in d
    assert d == _domain() and d != UnitCube(2)
    x = np.random.RandomState(0).rand(10, 2) * 2 - 1
    np.testing.assert_allclose(t.backward(t.forward(x)), x)
    np.testing.assert_allclose((~t).forward(t.forward(x)), x)


def test_linear_transform_backward_variance():       # reference test_transforms.py:82-99
    t = ~LinearTransform([2.0, 1.0], [1.2, 0.7])
    B = np.random.RandomState(1).rand(10, 2)
    np.testing.assert_allclose(t.backward_variance(B), B * np.array([4, 1]))
    t1, t2 = LinearTransform([2.0, 1.0], [1.2, 0.7]), LinearTransform([1.0, 1.0], [0, 0])
    t1.assign(t2)
    np.testing.assert_allclose(t1.A, t2.A)
    assert not LinearTransform(np.array([[1.0, 0.2], [0.0, 1.0]]), [0, 0]).is_diagonal()


def test_designs():
    d = _domain()
    X = RandomDesign(200, d).generate()
    assert X.shape == (200, 2) and X in d
    F = FactorialDesign(4, d).generate()
    assert F.shape == (16, 2) and F in d
    assert EmptyDesign(d).generate().shape == (0, 2)


def test_candidate_optimizer_finds_planted_optimum():        # reference test_optimizers.py:94-105
    d = _domain()
    opt = CandidateOptimizer(d, FactorialDesign(4, d).generate())
    assert opt.get_initial().shape == (0, 2) and not opt.gradient_enabled()
    opt.candidates = np.vstack((opt.candidates, np.zeros((1, 2))))
    result = opt.optimize(parabola2d)
    assert result.success and np.allclose(result.x, 0) and np.allclose(result.fun, 0) and result.nfev == 17
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        opt.set_initial([1, 1])
        assert len(w) == 1 and issubclass(w[-1].category, UserWarning)
    opt2 = CandidateOptimizer(d, FactorialDesign(4, d).generate())
    opt2.domain = UnitCube(2)
    np.testing.assert_allclose(opt2.candidates, FactorialDesign(4, UnitCube(2)).generate())


def test_mc_optimizer_first_occurrence_and_interrupt():
    d = _domain()
    opt = MCOptimizer(d, 200)
    r = opt.optimize(parabola2d)
    assert r.success and r.nfev == 200 and r.x.shape == (1, 2) and 0 < r.fun < 2
    r = opt.optimize(KeyboardRaiser(0, parabola2d))
    assert not r.success


def test_scipy_and_staged_optimizers():                        # reference test_optimizers.py:108-190
    d = _domain()
    opt = SciPyOptimizer(d, maxiter=10)
    assert opt.config == {'tol': None, 'method': 'L-BFGS-B', 'options': {'maxiter': 10, 'disp': False}}
    opt.set_initial([-1, -1])
    r = opt.optimize(parabola2d)
    assert r.success and r.nit <= 10 and r.nfev <= 20 and np.allclose(r.x, 0) and np.allclose(r.fun, 0)
    r = opt.optimize(KeyboardRaiser(2, parabola2d))
    assert not r.success and not np.allclose(r.x, 0)
    staged = StagedOptimizer([MCOptimizer(d, 5), MCOptimizer(d, 5), SciPyOptimizer(d, maxiter=10)])
    assert staged.gradient_enabled() is False
    r = staged.optimize(parabola2d)
    assert r.success and r.nstages == 3 and np.allclose(r.x, 0, atol=1e-3)
    r = staged.optimize(KeyboardRaiser(0, parabola2d))
    assert not r.success and r.nstages == 1 and r.nfev == 0


def test_multi_start_lock_step_equals_independent_runs():
    """the k polish runs advance in lock step on batched objective calls, and every run follows exactly
    the trajectory it would follow alone (same numbers in, same L-BFGS-B decisions out)."""
    d = _domain()
    calls = []

    def bumpy(X):
        calls.append(X.shape[0])
        f = np.sum(X ** 2 + 0.3 * np.sin(7.0 * X + 0.5), axis=1, keepdims=True)
        return f, 2 * X + 2.1 * np.cos(7.0 * X + 0.5)

    np.random.seed(11)
    opt = MultiStartOptimizer(d, 200, nstarts=6, maxiter=50)
    r = opt.optimize(bumpy)
    assert r.success and r.x.shape == (1, 2) and r.nstarts == 6
    assert calls[0] == 200 and max(calls[1:]) == 6 and len(calls) - 1 == r.nbatches
    assert r.nfev == 200 + r.polish_rows == sum(calls)
    assert r.nbatches == max(run.nfev for run in r.runs)               # lock step: as many batches as the longest run needs
    assert r.polish_rows == sum(run.nfev for run in r.runs)
    # the same starts, one at a time through the reference-style optimiser
    np.random.seed(11)
    starts = RandomDesign(200, d).generate()
    order = np.argsort(np.ravel(bumpy(starts)[0]), kind='stable')[:6]
    funs = []
    for i, x0 in enumerate(starts[order]):
        single = SciPyOptimizer(d, maxiter=50)
        single.set_initial(x0)
        rs = single.optimize(bumpy)
        np.testing.assert_array_equal(rs.x.ravel(), r.runs[i].x)       # bit-identical trajectory end points
        funs.append(float(np.ravel(rs.fun)[0]))
    assert float(np.ravel(r.fun)[0]) == min(funs)
    # errors inside a batched call reach the caller instead of dead-locking the other runs
    with pytest.raises(ZeroDivisionError):
        MultiStartOptimizer(d, 20, nstarts=3).optimize(lambda X: (bumpy(X)[0], bumpy(X)[1]) if X.shape[0] == 20 else 1 / 0)


def test_shard_bounds_and_pair_reduce():
    for n in (0, 1, 7, 8, 9, 1000):
        for ws in (1, 2, 3, 8):
            cover = []
            for r in range(ws):
                lo, hi = dist.shard_bounds(n, r, ws)
                assert 0 <= lo <= hi <= n
                cover.extend(range(lo, hi))
            assert cover == list(range(n))
    # max value wins; ties -> lowest global index; empty shards (-1) and NaN never win
    assert dist.reduce_pairs([1.0, 2.0, 2.0, 0.5], [3, 40, 12, 7]) == (2.0, 12)
    assert dist.reduce_pairs([0.0, -3.0], [-1, 5]) == (-3.0, 5)
    assert dist.reduce_pairs([float('nan'), -1.0], [0, 9]) == (-1.0, 9)
    assert dist.reduce_pairs([0.0], [-1]) == (0.0, -1)


class _FakeModel(object):
    class _Y(object):
        shape = (5, 1)
    Y = _Y()


def test_program_builder_postfix():
    m1, m2 = _FakeModel(), _FakeModel()
    b = ProgramBuilder({id(m1): 0, id(m2): 1})
    n1 = b.ei(m1, np.array([0.25]))
    n2 = b.pof(m2, 0.0)
    b.reduce('prod', [n1, n2])
    b.scale(None, 0.5)
    assert b.ops == [(_lib.OP_EI, 0, 0, 0), (_lib.OP_POF, 1, 0, 1), (_lib.OP_PROD, 2, 0, 0), (_lib.OP_SCALE, 0, 0, 2)]
    assert b.params == [0.25, 0.0, 0.5]


def test_device_candidate_stream_host_regeneration(libgpfo):
    """gpfo_random_rows (host, no GPU): rows are reproducible, shard-independent and inside the box."""
    from gpflowopt_b200._lib import random_rows
    lower, upper = np.array([-5.0, 0.0, 2.0]), np.array([10.0, 15.0, 2.5])
    a = random_rows(123, 0, 1000, lower, upper)
    b = np.vstack((random_rows(123, 0, 400, lower, upper), random_rows(123, 400, 600, lower, upper)))
    assert np.array_equal(a, b)
    assert np.all(a >= lower) and np.all(a < upper)
    assert not np.array_equal(a, random_rows(124, 0, 1000, lower, upper))
    u = (a - lower) / (upper - lower)
    assert abs(u.mean() - 0.5) < 0.02 and abs(u.var() - 1.0 / 12.0) < 0.01          # uniform first moments
    assert abs(np.corrcoef(u[:-1, 0], u[1:, 0])[0, 1]) < 0.08                        # no obvious serial correlation


def test_mes_gumbel_sampling_against_reference_generated_samples():
    """tests/golden/mes_golden.npz holds the y* samples the reference's own MinValueEntropySearch._setup (mes.py:63-94) draws
    for an analytic surrogate model and a fixed NumPy seed (tests/golden/gen_mes_golden.py).  The host-side restatement in
    gpflowopt_b200/acquisition/mes.py consumes the random stream in the same order and must reproduce them."""
    import importlib.util
    import os
    import types
    from gpflowopt_b200.acquisition.mes import MinValueEntropySearch
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("gen_mes_golden", os.path.join(here, "golden", "gen_mes_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)                          # only for surrogate_moments (no reference access at import)
    g = np.load(os.path.join(here, "golden", "mes_golden.npz"))
    for k in range(int(g["n_cases"])):
        X = g["case%d_X" % k]
        holder = lambda a: types.SimpleNamespace(value=a, shape=a.shape)
        model = types.SimpleNamespace(predict_f=gen.surrogate_moments, X=holder(X), Y=holder(gen.surrogate_moments(X)[0]))
        acq = object.__new__(MinValueEntropySearch)       # the class without its device-backed constructor
        acq._models, acq._parent, acq._engine, acq._needs_setup, acq.optimize_restarts = [model], None, None, False, 0
        acq.gridsize, acq.num_samples = int(g["case%d_gridsize" % k]), int(g["case%d_num_samples" % k])
        acq._domain = np.sum([ContinuousParameter("x%d" % i, l, u)
                              for i, (l, u) in enumerate(zip(g["case%d_lower" % k], g["case%d_upper" % k]))])
        np.random.seed(int(g["case%d_seed" % k]))
        acq._setup()
        np.testing.assert_allclose(acq.samples, g["case%d_samples" % k], rtol=1e-13, atol=0)


def test_bo_result_assembly_against_reference_generated_results():
    """tests/golden/bo_golden.npz: outputs of the reference's own BayesianOptimizer._create_bo_result (bo.py:156-203) on five
    evaluation records (single objective, Pareto set, constrained, nothing feasible, constraints only), generated by
    tests/golden/gen_bo_golden.py.  The restatement in gpflowopt_b200/bo.py must return the same arrays and flags."""
    import os
    import types
    from gpflowopt_b200.bo import _summarise
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bo_golden.npz"))
    for n in g["names"]:
        rec = types.SimpleNamespace(data=(g["%s_X" % n], g["%s_Y" % n]), feasible_data_index=lambda n=n: g["%s_feasible" % n],
                                    objective_indices=lambda n=n: g["%s_obj" % n], constraint_indices=lambda n=n: g["%s_con" % n])
        r = _summarise(rec, True, "OK")
        np.testing.assert_array_equal(np.atleast_2d(r.x), g["%s_x" % n])
        np.testing.assert_array_equal(np.atleast_2d(r.fun), g["%s_fun" % n])
        np.testing.assert_array_equal(np.atleast_2d(r.constraints), g["%s_constraints" % n])
        assert bool(r.success) == bool(g["%s_success" % n]) and str(r.message) == str(g["%s_message" % n])


def test_optimisers_against_reference_generated_results():
    """tests/golden/optim_golden.npz: results of the reference's own MCOptimizer / CandidateOptimizer / StagedOptimizer
    (gpflowopt/optim.py, row I) for seeded runs, a cross-row tie, a domain change and Ctrl-C at two places
    (tests/golden/gen_optim_golden.py).  The same script through gpflowopt_b200.optim must give identical results."""
    import importlib.util
    import os
    from gpflowopt_b200 import optim as product_optim
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("gen_optim_golden", os.path.join(here, "golden", "gen_optim_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)

    def make_domain(lo, up):
        return np.sum([ContinuousParameter("x%d" % i, l, u) for i, (l, u) in enumerate(zip(lo, up))])

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = gen.flatten(gen.run_all(product_optim, make_domain))
    want = np.load(os.path.join(here, "golden", "optim_golden.npz"))
    assert sorted(got) == sorted(want.files)
    for key in want.files:
        if key.endswith("_message"):
            assert str(got[key]) == str(want[key]), key
        else:
            np.testing.assert_array_equal(np.asarray(got[key]), np.asarray(want[key]), err_msg=key)


def test_aggregation_bookkeeping_against_reference_generated_records():
    """tests/golden/tree_golden.npz: the reference's own AcquisitionAggregation.set_data / _setup / constraint_indices /
    feasible_data_index run on recording stand-in operands (tests/golden/gen_tree_golden.py); the same script on
    gpflowopt_b200's class must leave the same call log and return the same arrays (including the reference's
    previous-operand shift of constraint columns for three or more operands)."""
    import importlib.util
    import os
    from gpflowopt_b200.acquisition.acquisition import Acquisition, AcquisitionAggregation
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("gen_tree_golden", os.path.join(here, "golden", "gen_tree_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)

    def make_aggregation(ops):
        agg = object.__new__(AcquisitionAggregation)
        agg.operands, agg._oper, agg._parent, agg._needs_setup, agg._engine = list(ops), 'sum', None, False, None
        return agg

    got = gen.script(make_aggregation)

    def make_leaf(tag):
        leaf = object.__new__(type("Leaf", (Acquisition,), {}))
        leaf.tag, leaf._parent, leaf._needs_setup, leaf._engine, leaf._models = tag, None, True, None, []
        return leaf
    got["operators"] = gen.operator_script(make_leaf)
    want = np.load(os.path.join(here, "golden", "tree_golden.npz"))
    assert sorted(got) == sorted(want.files)
    for key in want.files:
        np.testing.assert_array_equal(np.asarray(got[key]), np.asarray(want[key]), err_msg=key)


def test_leaf_bookkeeping_against_reference_generated_records():
    """tests/golden/leaf_golden.npz: the reference's own Acquisition._optimize_models (restart protocol: first run from the
    current state, the others randomised, Cholesky failures skipped, best run -- first on ties -- restored, all failing
    raises) and set_data (per-model column blocks, setup flag) run on recording stand-in models
    (tests/golden/gen_leaf_golden.py); gpflowopt_b200's class must leave the same call log."""
    import importlib.util
    import os
    from gpflowopt_b200.acquisition.acquisition import Acquisition
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("gen_leaf_golden", os.path.join(here, "golden", "gen_leaf_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)

    def make_leaf(models, restarts):
        leaf = object.__new__(Acquisition)
        leaf._models, leaf.optimize_restarts, leaf._needs_setup, leaf._parent, leaf._engine = models, restarts, True, None, None
        return leaf

    failure = lambda message: _lib.NotPositiveDefiniteError(_lib.ERR_NOT_SPD, message)
    got = gen.script(make_leaf, failure, None)
    want = np.load(os.path.join(here, "golden", "leaf_golden.npz"))
    assert sorted(got) == sorted(want.files)
    for key in want.files:
        np.testing.assert_array_equal(np.asarray(got[key]), np.asarray(want[key]), err_msg=key)


def test_bo_loop_against_reference_generated_record():
    """tests/golden/boloop_golden.npz: the reference's own BayesianOptimizer.optimize (initial design, then n_iter times
    model callback -> inner optimiser on the negated acquisition -> objective evaluation -> set_data) run on recording
    stand-ins (tests/golden/gen_boloop_golden.py); gpflowopt_b200's loop must leave the same call log and result."""
    import importlib.util
    import os
    from gpflowopt_b200.bo import BayesianOptimizer
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("gen_boloop_golden", os.path.join(here, "golden", "gen_boloop_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)

    def make_domain(lo, up):
        return np.sum([ContinuousParameter("x%d" % i, l, u) for i, (l, u) in enumerate(zip(lo, up))])

    def make_bo(dom, acq, inner, callback, initial):
        bo = object.__new__(BayesianOptimizer)
        bo._domain, bo._initial, bo._wrapper_args = dom, initial, dict(exclude_gradient=True)
        bo.acquisition, bo.optimizer, bo._model_callback, bo.verbose, bo._scaling = acq, inner, callback, False, False
        return bo

    got = gen.script(make_bo, make_domain)
    want = np.load(os.path.join(here, "golden", "boloop_golden.npz"))
    assert sorted(got) == sorted(want.files)
    for key in want.files:
        np.testing.assert_array_equal(np.asarray(got[key]), np.asarray(want[key]), err_msg=key)
