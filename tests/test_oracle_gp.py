"""Anchors the NumPy restatement of the GPflow-0.5 / TF-1 arithmetic (PARITY UNPINNED at 1e-9 by the reference
itself): reference fixtures at their own tolerances + 50-digit mpmath truth on small cases + internal identities."""
import os

import mpmath as mp
import numpy as np
import pytest

from oracle import acq as oacq, gp as ogp, pareto as opareto
from gpflowopt_b200 import synthetic
from helpers import GOLDEN, fit_mle, oracle_gp

REF_DATA = os.path.join(GOLDEN, "vlmop.npz")


def test_backward_variance_known_answer():
    """reference testing/unit/test_transforms.py:82-99: ~LinearTransform([2,1],[1.2,0.7]) scales variances by [4,1]."""
    t = ogp.Linear([2.0, 1.0], [1.2, 0.7]).invert()
    B = np.random.RandomState(0).rand(10, 2)
    np.testing.assert_allclose(t.backward_variance_literal(B), B * np.array([4, 1]))
    np.testing.assert_allclose(t.backward_variance(B), B * np.array([4, 1]))


def test_forward_backward_roundtrip():
    """reference testing/unit/test_transforms.py:34-54."""
    t = ogp.Linear([2.0, 3.5], [1.2, 0.7])
    x = np.random.RandomState(1).rand(10, 2)
    y = t.forward(x)
    np.testing.assert_allclose(t.backward(y), x)
    np.testing.assert_allclose(t.invert().forward(y), x)


def test_domain_transform():
    """reference gpflowopt/domain.py:89-93: box -> box."""
    t = ogp.domain_transform([-5.0, 0.0], [10.0, 15.0], [0.0, 0.0], [1.0, 1.0])
    np.testing.assert_allclose(t.forward(np.array([[-5.0, 0.0], [10.0, 15.0], [2.5, 7.5]])),
                               [[0, 0], [1, 1], [0.5, 0.5]], atol=1e-15)


def test_literal_and_linear_variance_forms_agree():
    X, Y = synthetic.training_set(40, 3)
    g = oracle_gp(X * 15 - 5, Y * 3 + 2, ogp.MATERN52, synthetic.hypers(3), lower=[-5] * 3, upper=[10] * 3,
                  normalize=True)
    Xc = synthetic.candidates(50, 3, lower=[-5] * 3, upper=[10] * 3)
    m1, v1 = g.predict(Xc, literal_variance=True)
    m2, v2 = g.predict(Xc)
    np.testing.assert_allclose(v1, v2, rtol=1e-13)
    np.testing.assert_array_equal(m1, m2)


def test_scaled_and_unscaled_predictions_agree():
    """reference testing/unit/test_datascaler.py:70-96 (atol 1e-3 there): scaling inputs to the unit cube with
    proportionally scaled lengthscales leaves predictions unchanged."""
    X, Y = synthetic.training_set(30, 2)
    lo, up = np.array([-1.0, -1.0]), np.array([1.0, 1.0])
    Xd = lo + X * (up - lo)
    h = synthetic.hypers(2)
    g_scaled = oracle_gp(Xd, Y, ogp.RBF, h, lower=lo, upper=up)
    h2 = dict(h, lengthscales=h["lengthscales"] * (up - lo))
    g_plain = oracle_gp(Xd, Y, ogp.RBF, h2)
    Xc = lo + synthetic.candidates(20, 2) * (up - lo)
    for a, b in zip(g_scaled.predict(Xc), g_plain.predict(Xc)):
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-12)


def _mp_predict(g, Xc):
    """50-digit GP prediction + EI with exact (non-expanded) distances."""
    mp.mp.dps = 50
    Xs = [[mp.mpf(v) for v in row] for row in g.Xs]
    ell = [mp.mpf(v) for v in np.broadcast_to(g.lengthscales, (g.Xs.shape[1],))]
    var, noise = mp.mpf(g.variance), mp.mpf(g.noise_variance)

    def k(a, b):
        r2 = sum(((ai - bi) / li) ** 2 for ai, bi, li in zip(a, b, ell))
        if g.kind == ogp.RBF:
            return var * mp.exp(-r2 / 2)
        r = mp.sqrt(r2 + mp.mpf('1e-12'))
        if g.kind == ogp.MATERN52:
            return var * (1 + mp.sqrt(5) * r + mp.mpf(5) / 3 * r * r) * mp.exp(-mp.sqrt(5) * r)
        return var * (1 + mp.sqrt(3) * r) * mp.exp(-mp.sqrt(3) * r)

    M = len(Xs)
    K = mp.matrix(M, M)
    for i in range(M):
        for j in range(M):
            K[i, j] = k(Xs[i], Xs[j]) + (noise if i == j else 0)
    y = mp.matrix([mp.mpf(v) for v in g.Ys[:, 0]])
    Kinv_y = mp.lu_solve(K, y)
    a_in, b_in = np.diag(g.input_transform.A), g.input_transform.b
    a_out, b_out = mp.mpf(np.diag(g.output_transform.A)[0]), mp.mpf(g.output_transform.b[0])
    means, vars_ = [], []
    for row in Xc:
        xs = [mp.mpf(v) * mp.mpf(a) + mp.mpf(b) for v, a, b in zip(row, a_in, b_in)]
        kx = mp.matrix([k(xi, xs) for xi in Xs])
        mu = sum(kx[i] * Kinv_y[i] for i in range(M))
        w = mp.lu_solve(K, kx)
        v = var - sum(kx[i] * w[i] for i in range(M))
        means.append((mu - b_out) / a_out)
        vars_.append(v / (a_out * a_out))
    return means, vars_


@pytest.mark.parametrize("kind", [ogp.RBF, ogp.MATERN52, ogp.MATERN32])
def test_against_mpmath_truth(kind):
    """Bounds the restatement against exact arithmetic: mean / var / EI / PoF within 1e-9 of the batch scale."""
    M, D, N = 24, 2, 12
    X, Y = synthetic.training_set(M, D)
    lo, up = np.array([-5.0, 0.0]), np.array([10.0, 15.0])
    Xd = lo + X * (up - lo)
    g = oracle_gp(Xd, Y * 2 + 1, kind, synthetic.hypers(D), lower=lo, upper=up, normalize=True)
    Xc = lo + synthetic.candidates(N, D) * (up - lo)
    mean, var = g.predict(Xc)
    m_mp, v_mp = _mp_predict(g, Xc)
    m_ref = np.array([float(v) for v in m_mp])
    v_ref = np.array([float(v) for v in v_mp])
    assert np.max(np.abs(mean[:, 0] - m_ref)) <= 1e-9 * np.max(np.abs(m_ref))
    prior_var = g.variance / np.diag(g.output_transform.A)[0] ** 2
    assert np.max(np.abs(var[:, 0] - v_ref)) <= 1e-9 * prior_var
    fmin = float(np.min(g.predict(Xd)[0]))
    ei = oacq.Leaf('ei', g, fmin).value(Xc)[:, 0]
    pof = oacq.Leaf('pof', g, 0.3).value(Xc)[:, 0]
    ei_ref, pof_ref = [], []
    for m, v in zip(m_mp, v_mp):
        v = max(v, mp.mpf('1e-6'))
        s = mp.sqrt(v)
        z = (mp.mpf(fmin) - m) / s
        ei_ref.append(float((mp.mpf(fmin) - m) * mp.ncdf(z) + s * mp.npdf(z)))
        pof_ref.append(float(mp.ncdf((mp.mpf('0.3') - m) / s)))
    assert np.max(np.abs(ei - np.array(ei_ref))) <= 1e-9 * np.max(np.abs(ei_ref))
    assert np.max(np.abs(pof - np.array(pof_ref))) <= 1e-9


def test_ndtr_matches_mpmath_in_all_branches():
    t = np.array([-38.0, -12.5, -3.0, -1.0, -0.999, -0.3, 0.0, 0.4, 0.999, 1.0, 1.0001, 2.5, 9.0, 40.0, -np.inf, np.inf])
    got = oacq.ndtr(t)
    mp.mp.dps = 50
    for x, g_ in zip(t, got):
        ref = float(mp.ncdf(mp.mpf(x))) if np.isfinite(x) else (0.0 if x < 0 else 1.0)
        assert abs(g_ - ref) <= 2e-15 * max(ref, 1e-300) + 1e-320, (x, g_, ref)   # erfc: a few ulp


def test_hvpoi_reference_fixture():
    """reference testing/unit/test_implementations.py:184-187 with testing/data/vlmop.npz (decimal=2): Matern32 GPRs
    whose hyper-parameters come from a maximum-likelihood fit (GPflow model.optimize(); restated in helpers.fit_mle)."""
    d = np.load(REF_DATA)
    gps = []
    for r in range(2):
        h = fit_mle(d['X'], d['Y'][:, [r]], ogp.MATERN32)
        gps.append(oracle_gp(d['X'], d['Y'][:, [r]], ogp.MATERN32, h))
    F = np.hstack([g.predict(d['X'])[0] for g in gps])
    p = opareto.Pareto(np.zeros((1, 2)))
    p.update(F)
    hv = oacq.HVPoI(gps, p.front, p.lb, p.ub)
    scores = hv.value(d['candidates'])
    np.testing.assert_almost_equal(scores, d['scores'], decimal=2)


def test_aggregation_algebra():
    """reference testing/unit/test_acquisition.py:170-191: Sum of two equal EIs = 2 EI, Product = EI^2."""
    X, Y = synthetic.training_set(16, 2)
    g = oracle_gp(X, Y, ogp.RBF, synthetic.hypers(2))
    fmin = float(np.min(g.predict(X)[0]))
    Xc = synthetic.candidates(16, 2)
    ei = oacq.Leaf('ei', g, fmin)
    np.testing.assert_allclose(oacq.Agg('sum', [ei, ei]).value(Xc), 2 * ei.value(Xc), rtol=1e-14)
    np.testing.assert_allclose(oacq.Agg('prod', [ei, ei]).value(Xc), ei.value(Xc) ** 2, rtol=1e-14)


def test_lcb_sigma_zero_is_the_mean():
    """reference testing/unit/test_implementations.py:162-168."""
    X, Y = synthetic.training_set(25, 2)
    g = oracle_gp(X, Y, ogp.RBF, synthetic.hypers(2))
    Xc = synthetic.candidates(200, 2)
    np.testing.assert_allclose(oacq.Leaf('lcb', g, 0.0).value(Xc), g.predict(Xc)[0], rtol=1e-7)
    assert np.all(oacq.Leaf('lcb', g, 3.2).value(Xc) < g.predict(Xc)[0])


def _autograd_reference(node, Xc):
    """torch.autograd over the same op graph (what tf.gradients does in the reference)."""
    torch = pytest.importorskip("torch")
    g = node.gp
    L, V = (torch.from_numpy(a) for a in g.factor())
    Xs = torch.from_numpy(g.Xs)
    x = torch.tensor(Xc, dtype=torch.float64, requires_grad=True)
    a_in, b_in = torch.from_numpy(np.diag(g.input_transform.A).copy()), torch.from_numpy(g.input_transform.b)
    ell = torch.from_numpy(np.broadcast_to(g.lengthscales, (Xc.shape[1],)).copy())
    xn = (x * a_in + b_in) / ell
    Xl = Xs / ell
    r2 = -2.0 * Xl @ xn.T + (Xl ** 2).sum(1)[:, None] + (xn ** 2).sum(1)[None, :]
    if g.kind == ogp.RBF:
        Kx = g.variance * torch.exp(-r2 / 2)
    else:
        r = torch.sqrt(r2 + 1e-12)
        Kx = g.variance * (1 + np.sqrt(5.0) * r + 5.0 / 3.0 * r ** 2) * torch.exp(-np.sqrt(5.0) * r)
    A = torch.linalg.solve_triangular(L, Kx, upper=False)
    a_out, b_out = float(np.diag(g.output_transform.A)[0]), float(g.output_transform.b[0])
    mean = ((A.T @ V)[:, 0] - b_out) / a_out
    var = (g.variance - (A ** 2).sum(0)) / a_out ** 2
    p = getattr(node, 'param', None)
    if node.kind == 'mes':
        gamma = (mean[:, None] - torch.from_numpy(node.samples)[None, :]) / torch.sqrt(var)[:, None]
        pdf = torch.exp(-0.5 * gamma ** 2 - 0.5 * np.log(2 * np.pi))
        cdf = 0.5 * (1 + torch.erf(gamma / np.sqrt(2.0)))
        val = (gamma * pdf / (2.0 * cdf) - torch.special.log_ndtr(gamma)).sum(1) / node.samples.size
    elif node.kind == 'lcb':
        val = mean - p * torch.sqrt(torch.clamp(var, min=0.0))
    else:
        v = torch.clamp(var, min=1e-6)
        s = torch.sqrt(v)
        z = (p - mean) / s
        Phi = 0.5 * (1 + torch.erf(z / np.sqrt(2.0)))
        if node.kind == 'ei':
            val = (p - mean) * Phi + v * torch.exp(-0.5 * z ** 2 - 0.5 * np.log(2 * np.pi) - torch.log(s))
        else:
            val = Phi
    val.sum().backward()
    return val.detach().numpy(), x.grad.numpy()


@pytest.mark.parametrize("kind", [ogp.RBF, ogp.MATERN52])
@pytest.mark.parametrize("leaf", ['ei', 'poi', 'pof', 'lcb'])
def test_closed_form_gradient_matches_autograd(kind, leaf):
    X, Y = synthetic.training_set(30, 3)
    lo, up = np.full(3, -2.0), np.full(3, 3.0)
    Xd = lo + X * (up - lo)
    g = oracle_gp(Xd, 2 * Y - 1, kind, synthetic.hypers(3), lower=lo, upper=up, normalize=True)
    Xc = lo + synthetic.candidates(25, 3) * (up - lo)
    node = oacq.Leaf(leaf, g, {'ei': -0.7, 'poi': -0.7, 'pof': 0.1, 'lcb': 2.0}[leaf])
    val, grad = node.value_and_grad(Xc)
    v_ref, g_ref = _autograd_reference(node, Xc)
    np.testing.assert_allclose(val[:, 0], v_ref, rtol=1e-10, atol=1e-14)
    assert np.max(np.abs(grad - g_ref)) <= 1e-9 * np.max(np.abs(g_ref))
    # and the value returned with gradients equals evaluate() (reference test_implementations.py:19-33)
    np.testing.assert_array_equal(val, node.value(Xc))


def test_log_ndtr_branches_against_mpmath():
    """TF1 log_ndtr segments (float64: series below -20, log(ndtr) up to 8, -ndtr(-x) above) against exact values; the
    3-term series itself is only good to 105 / x^8 (4e-9 at the switch point), which bounds what the reference computes."""
    mp.mp.dps = 50
    x = np.array([-40.0, -25.0, -20.5, -20.0, -19.99, -8.0, -1.0, 0.0, 0.7, 3.0, 7.99, 8.0, 8.01, 12.0])
    got, dgot = oacq.log_ndtr(x), oacq.dlog_ndtr(x)
    for xi, g_, d_ in zip(x, got, dgot):
        ref = float(mp.log(mp.ncdf(mp.mpf(xi))))
        dref = float(mp.npdf(mp.mpf(xi)) / mp.ncdf(mp.mpf(xi)))
        tol = 1.1 * 105.0 / xi ** 8 if xi <= -20.0 else 1e-13 * abs(ref) + 2.3e-16    # log(1 - q): one ulp of 1
        assert abs(g_ - ref) <= tol, (xi, g_, ref)
        dtol = 1e-6 * abs(dref) if xi <= -20.0 else 1e-12 * abs(dref) + 1e-30
        assert abs(d_ - dref) <= dtol, (xi, d_, dref)


@pytest.mark.parametrize("kind", [ogp.RBF, ogp.MATERN52])
def test_mes_value_against_mpmath_and_gradient_against_autograd(kind):
    """mes.py:96-102 restated: value against 50-digit arithmetic on the mpmath posterior, gradient against autograd."""
    M, D, N = 24, 2, 12
    X, Y = synthetic.training_set(M, D)
    lo, up = np.array([-5.0, 0.0]), np.array([10.0, 15.0])
    Xd = lo + X * (up - lo)
    g = oracle_gp(Xd, Y * 2 + 1, kind, synthetic.hypers(D), lower=lo, upper=up, normalize=True)
    Xc = lo + synthetic.candidates(N, D) * (up - lo)
    fmin = float(np.min(g.predict(Xd)[0]))
    samples = fmin - np.array([0.05, 0.2, 0.6, 1.5, 4.0])
    node = oacq.MES(g, samples)
    val = node.value(Xc)[:, 0]
    m_mp, v_mp = _mp_predict(g, Xc)
    ref = []
    for m, v in zip(m_mp, v_mp):
        acc = mp.mpf(0)
        for y in samples:
            gam = (m - mp.mpf(float(y))) / mp.sqrt(v)
            acc += gam * mp.npdf(gam) / (2 * mp.ncdf(gam)) - mp.log(mp.ncdf(gam))
        ref.append(float(acc / len(samples)))
    assert np.max(np.abs(val - np.array(ref))) <= 1e-9 * np.max(np.abs(ref))
    v2, grad = node.value_and_grad(Xc)
    np.testing.assert_array_equal(v2[:, 0], val)
    if kind in (ogp.RBF, ogp.MATERN52):
        v_ref, g_ref = _autograd_reference(node, Xc)
        np.testing.assert_allclose(val, v_ref, rtol=1e-10, atol=1e-14)
        assert np.max(np.abs(grad - g_ref)) <= 1e-9 * np.max(np.abs(g_ref))


def test_product_and_hvpoi_gradients_by_finite_differences():
    X, Y2 = synthetic.training_set(20, 2, constraint=True)
    h = synthetic.hypers(2)
    g1 = oracle_gp(X, Y2[:, [0]], ogp.MATERN52, h)
    g2 = oracle_gp(X, Y2[:, [1]], ogp.RBF, h)
    fmin = float(np.min(g1.predict(X)[0]))
    prod = oacq.Agg('prod', [oacq.Leaf('ei', g1, fmin), oacq.Leaf('pof', g2, 0.0)])
    F = np.hstack([g.predict(X)[0] for g in (g1, g2)])
    p = opareto.Pareto(np.zeros((1, 2)))
    p.update(F)
    hv = oacq.HVPoI([g1, g2], p.front, p.lb, p.ub)
    Xc = synthetic.candidates(6, 2)
    for node in (prod, hv):
        val, grad = node.value_and_grad(Xc)
        np.testing.assert_allclose(val, node.value(Xc), rtol=1e-13)
        eps = 1e-6
        for d in range(2):
            e = np.zeros(2)
            e[d] = eps
            fd = (node.value(Xc + e) - node.value(Xc - e))[:, 0] / (2 * eps)
            np.testing.assert_allclose(grad[:, d], fd, rtol=2e-5, atol=1e-9 * np.max(np.abs(grad)))


def test_mc_optimize_first_occurrence():
    """reference gpflowopt/optim.py:155-164: np.argmin returns the first minimum."""
    X, Y = synthetic.training_set(16, 2)
    g = oracle_gp(X, Y, ogp.RBF, synthetic.hypers(2))
    ei = oacq.Leaf('ei', g, float(np.min(g.predict(X)[0])))
    Xc = synthetic.candidates(50, 2)
    Xc[37] = Xc[5]                                  # exact duplicate of an earlier row
    scores = ei.value(Xc)
    Xc2 = Xc.copy()
    best = int(np.argmax(scores))
    Xc2[45] = Xc2[best]
    x, f, idx, _ = oacq.mc_optimize(ei, Xc2)
    assert idx == min(best, 45) and f[0, 0] == -scores[best, 0]


def test_domain_maps_and_membership_against_reference_generated_vectors():
    """tests/golden/domain_golden.npz holds outputs of the reference's own gpflowopt/domain.py (Domain.__rshift__ coefficients
    and Domain.__contains__ verdicts, generated by tests/golden/gen_domain_golden.py): the oracle's and the product's host
    classes must reproduce them bit for bit -- this pins the input transform of SURVEY 8a row B."""
    import os
    from helpers import GOLDEN
    from gpflowopt_b200.domain import ContinuousParameter
    g = np.load(os.path.join(GOLDEN, "domain_golden.npz"))

    def box(lo, up):
        return np.sum([ContinuousParameter("p%d" % i, l, u) for i, (l, u) in enumerate(zip(lo, up))])

    for k in range(int(g["n_pairs"])):
        lo1, up1, lo2, up2 = g["pair%d_bounds" % k]
        A, b = g["pair%d_A" % k], g["pair%d_b" % k]
        t_oracle = ogp.domain_transform(lo1, up1, lo2, up2)
        np.testing.assert_array_equal(t_oracle.A, A)
        np.testing.assert_array_equal(t_oracle.b, b)
        t_product = box(lo1, up1) >> box(lo2, up2)
        np.testing.assert_array_equal(t_product.A, A)
        np.testing.assert_array_equal(t_product.b, b)
    dom = box(*g["contains_bounds"])
    verdicts = [bool(p[None, :] in dom) for p in g["contains_probes"]]
    assert verdicts == [bool(v) for v in g["contains_rows"]]
    assert (g["contains_probes"][:3] in dom) == bool(g["contains_all"])
    assert (np.zeros((2, 3)) in dom) == bool(g["contains_wrong_width"])


def test_acquisition_compositions_against_reference_executed_graphs():
    """tests/golden/acq_golden.npz: the reference's own ``build_acquisition`` methods (EI, PoI, PoF, LCB, MES, HVPoI) executed
    on fixed predictive moments with NumPy stand-ins for the TensorFlow primitives (tests/golden/gen_acq_golden.py).  The
    oracle's compositions must give the same numbers: this pins clamps, operand order, gather indices and reductions at the
    GPflowOpt level (the TF-internal ndtr / prob forms are the oracle's own on both sides)."""
    import os
    from helpers import GOLDEN
    g = np.load(os.path.join(GOLDEN, "acq_golden.npz"))
    mean, var = g["mean"], g["var"]
    leaf = lambda kind, p: oacq.Leaf(kind, None, p)._value(mean, var)
    np.testing.assert_array_equal(leaf('ei', g["fmin"][0]), g["ei"])
    np.testing.assert_array_equal(leaf('poi', g["fmin"][0]), g["poi"])
    np.testing.assert_array_equal(leaf('pof', float(g["threshold"])), g["pof"])
    np.testing.assert_array_equal(leaf('lcb', float(g["sigma"])), g["lcb"])
    with np.errstate(all="ignore"):
        mes = oacq.MES(None, g["samples"])._value(mean, var)
    np.testing.assert_array_equal(np.isnan(mes), np.isnan(g["mes"]))
    ok = np.isfinite(g["mes"])
    np.testing.assert_array_equal(mes[ok], g["mes"][ok])
    hv = oacq.HVPoI([], g["hv_front"], g["hv_lb"], g["hv_ub"], reference=g["hv_reference"])
    with np.errstate(invalid="ignore"):
        got = hv._value_from_moments(g["hv_mean"], np.maximum(g["hv_var"], oacq.STABILITY))
    np.testing.assert_allclose(got, g["hvpoi"], rtol=1e-13, atol=1e-300)
    assert np.max(g["hvpoi"]) > 0


def test_transforms_against_reference_executed_graph_builders():
    """tests/golden/transform_golden.npz: the reference's own LinearTransform.build_forward / build_backward /
    build_backward_variance, ``~LinearTransform(std, mean)`` and ``Domain >> UnitCube`` executed on fixed data
    (tests/golden/gen_transform_golden.py).  The oracle's literal restatement reproduces them bit for bit; the O(N) variance
    form and the product's host class agree to rounding (one division instead of two triangular solves)."""
    import os
    from helpers import GOLDEN
    from gpflowopt_b200.transforms import LinearTransform
    g = np.load(os.path.join(GOLDEN, "transform_golden.npz"))
    for k in range(int(g["n_cases"])):
        c = lambda name: g["c%d_%s" % (k, name)]
        t_in = ogp.domain_transform(c("lower"), c("upper"), np.zeros(c("lower").size), np.ones(c("lower").size))
        np.testing.assert_array_equal(t_in.A, c("in_A"))
        np.testing.assert_array_equal(t_in.b, c("in_b"))
        np.testing.assert_array_equal(t_in.forward(c("X")), c("X_scaled"))
        np.testing.assert_array_equal(t_in.backward(c("X_scaled")), c("X_back"))
        Y = c("Y")
        t_out = ogp.Linear(Y.std(axis=0), Y.mean(axis=0)).invert()
        np.testing.assert_array_equal(t_out.A, c("out_A"))
        np.testing.assert_array_equal(t_out.b, c("out_b"))
        np.testing.assert_array_equal(t_out.forward(Y), c("Y_scaled"))
        np.testing.assert_array_equal(t_out.backward(c("f_scaled")), c("mean"))
        np.testing.assert_array_equal(t_out.backward_variance_literal(c("v_scaled")), c("var"))
        np.testing.assert_allclose(t_out.backward_variance(c("v_scaled")), c("var"), rtol=4e-16)
        # the product's host-side class (diagonal maps; its forward is what feeds gpfo_gp_set)
        p_out = ~LinearTransform(Y.std(axis=0), Y.mean(axis=0))
        np.testing.assert_array_equal(p_out.A, c("out_A"))
        np.testing.assert_array_equal(p_out.b, c("out_b"))
        np.testing.assert_array_equal(p_out.forward(Y), c("Y_scaled"))
        np.testing.assert_allclose(p_out.backward(c("f_scaled")), c("mean"), rtol=1e-15)
        np.testing.assert_allclose(p_out.backward_variance(c("v_scaled")), c("var"), rtol=1e-15)
