"""GPU parity tests proper (-m gpu): the CUDA path, called through the C-ABI, against the CPU oracle on the same
seeded inputs.  Tolerance (written here, from BASELINE.json north_star "rel. 1e-9 (float64)"): scale-relative 1e-9,
see helpers.RTOL; argmax index identical except on near-ties within that tolerance."""
import numpy as np
import pytest

from gpflowopt_b200 import _lib, synthetic
from oracle import acq as oacq, gp as ogp, pareto as opareto
from helpers import RTOL, assert_parity, near_tie, oracle_gp, upload_oracle_gp

pytestmark = pytest.mark.gpu

KINDS = {"rbf": ogp.RBF, "matern52": ogp.MATERN52, "matern32": ogp.MATERN32}

# kernel variants of the contraction (csrc/contract.cu): production ones first; 10 / 13 / 14 (rejected experiments) exist only
# in libraries built with GPFO_BUILD_EXPERIMENTS=1 and are skipped otherwise
PRODUCTION_VARIANTS = [20, 21, 12, 11, 4, 5, 6, 19]
EXPERIMENT_VARIANTS = [14, 13, 10]


def _set_variant_or_skip(ctx, variant):
    try:
        ctx.set_variant(variant)
    except _lib.GpfoError:
        if variant in EXPERIMENT_VARIANTS:
            pytest.skip("variant %d needs a GPFO_BUILD_EXPERIMENTS=1 build" % variant)
        raise


def _model(M, D, kind, noise=1e-2, scaled=False, seed=1234, R=1):
    X, Y = synthetic.training_set(M, D, seed=seed, constraint=(R == 2))
    h = synthetic.hypers(D, noise)
    if scaled:                     # domain [-5, 10]^D with input scaling and output normalisation
        lo, up = np.full(D, -5.0), np.full(D, 10.0)
        return oracle_gp(lo + X * (up - lo), 3.0 * Y + 2.0, KINDS[kind], h, lower=lo, upper=up, normalize=True), lo, up
    return oracle_gp(X, Y, KINDS[kind], h), np.zeros(D), np.ones(D)


def _cands(N, D, lo, up, seed=4321):
    return lo + synthetic.candidates(N, D, seed=seed) * (up - lo)


@pytest.mark.parametrize("M", [1, 7, 64, 65, 200, 513])
def test_factor_matches_lapack(ctx, M):
    g, _, _ = _model(M, 3, "matern52")
    upload_oracle_gp(ctx, 0, g)
    L, Linv, V = ctx.gp_get_factor(0, M, 1)
    Lo, Vo = g.factor()
    assert_parity(L, Lo, tol=1e-12, what="L")
    assert_parity(V, Vo, tol=1e-11, what="V")
    assert np.max(np.abs(Linv @ Lo - np.eye(M))) < 1e-10
    assert np.all(np.triu(Linv, 1) == 0.0) and np.all(np.triu(L, 1) == 0.0)


def test_not_positive_definite_raises(ctx):
    X = np.zeros((8, 2))                                   # 8 identical points, no noise: singular K
    with pytest.raises(_lib.NotPositiveDefiniteError):
        ctx.gp_set(0, X, np.zeros((8, 1)), _lib.KERN_RBF, [1.0, 1.0], 1.0, 0.0)


@pytest.mark.parametrize("kind", ["rbf", "matern52", "matern32"])
@pytest.mark.parametrize("M,D,scaled", [(20, 2, False), (100, 3, True), (515, 8, False), (1100, 16, True)])
def test_predict_parity(ctx, kind, M, D, scaled):
    g, lo, up = _model(M, D, kind, scaled=scaled)
    upload_oracle_gp(ctx, 0, g)
    for N in (1, 9, 1000):
        Xc = _cands(N, D, lo, up)
        mean, var = ctx.predict(0, Xc, 1)
        mo, vo = g.predict(Xc)
        assert_parity(mean, mo, what="mean M=%d N=%d" % (M, N))
        prior = g.variance / np.diag(g.output_transform.A)[0] ** 2
        assert_parity(var, vo, scale=prior, what="var M=%d N=%d" % (M, N))


def test_multi_output_gp(ctx):
    """One GP with two output columns (shared covariance): mean per column, variance tiled (GPflow gpr.py)."""
    g, lo, up = _model(150, 4, "matern52", R=2)
    upload_oracle_gp(ctx, 0, g)
    Xc = _cands(333, 4, lo, up)
    mean, var = ctx.predict(0, Xc, 2)
    mo, vo = g.predict(Xc)
    assert_parity(mean, mo, what="mean")
    assert_parity(var, vo, scale=g.variance, what="var")


@pytest.mark.parametrize("variant", PRODUCTION_VARIANTS + EXPERIMENT_VARIANTS)
@pytest.mark.parametrize("leaf,param", [("ei", None), ("poi", None), ("pof", 0.15), ("lcb", 2.0)])
def test_leaf_acquisitions(ctx, variant, leaf, param):
    g, lo, up = _model(300, 5, "matern52", scaled=True)
    upload_oracle_gp(ctx, 0, g)
    if param is None:
        param = float(np.min(g.predict(g.input_transform.backward(g.Xs))[0]))     # fmin over the training inputs
    op = {"ei": _lib.OP_EI, "poi": _lib.OP_POI, "pof": _lib.OP_POF, "lcb": _lib.OP_LCB}[leaf]
    ctx.program_set([[op, 0, 0, 0]], [param])
    _set_variant_or_skip(ctx, variant)
    try:
        for N in (1, 31, 33, 2000):
            Xc = _cands(N, 5, lo, up)
            vals, _, bv, bi = ctx.eval(Xc)
            ref = oacq.Leaf(leaf, g, param).value(Xc)
            assert_parity(vals, ref, what="%s N=%d variant %d" % (leaf, N, variant))
            assert bv == vals[bi, 0] and bi == int(np.argmax(vals))           # fused argmax == argmax of its own scores
            assert bi == int(np.argmax(ref)) or near_tie(ref, bi, int(np.argmax(ref)))
    finally:
        ctx.set_variant(0)


def test_sum_product_scale_programs(ctx):
    """EI x PoF (constrained BO), nested EI * (PoF + EI), Sum, and the MCMC mean (SUM then SCALE 1/n)."""
    D, N = 3, 777
    g1, lo, up = _model(120, D, "matern52")
    g2, _, _ = _model(90, D, "rbf", seed=77)
    g3, _, _ = _model(60, D, "matern32", seed=99, noise=1e-3)
    for i, g in enumerate((g1, g2, g3)):
        upload_oracle_gp(ctx, i, g, slot0=i)
    ctx.gp_count_set(3)
    Xc = _cands(N, D, lo, up)
    f1, f3 = -0.4, 0.2
    ei1, pof2, ei3 = oacq.Leaf('ei', g1, f1), oacq.Leaf('pof', g2, 0.0), oacq.Leaf('ei', g3, f3)
    cases = [
        ([[_lib.OP_EI, 0, 0, 0], [_lib.OP_POF, 1, 0, 1], [_lib.OP_PROD, 2, 0, 0]], [f1, 0.0],
         oacq.Agg('prod', [ei1, pof2])),
        ([[_lib.OP_EI, 0, 0, 0], [_lib.OP_POF, 1, 0, 1], [_lib.OP_EI, 2, 0, 2], [_lib.OP_SUM, 2, 0, 0],
          [_lib.OP_PROD, 2, 0, 0]], [f1, 0.0, f3], oacq.Agg('prod', [ei1, oacq.Agg('sum', [pof2, ei3])])),
        ([[_lib.OP_EI, 0, 0, 0], [_lib.OP_POF, 1, 0, 1], [_lib.OP_EI, 2, 0, 2], [_lib.OP_SUM, 3, 0, 0]],
         [f1, 0.0, f3], oacq.Agg('sum', [ei1, pof2, ei3])),
        ([[_lib.OP_EI, 0, 0, 0], [_lib.OP_EI, 2, 0, 1], [_lib.OP_SUM, 2, 0, 0], [_lib.OP_SCALE, 0, 0, 2]],
         [f1, f3, 0.5], oacq.Agg('sum', [ei1, ei3], scale=0.5)),
    ]
    try:
        for ops, params, node in cases:
            ctx.program_set(ops, params)
            vals, _, bv, bi = ctx.eval(Xc)
            ref = node.value(Xc)
            assert_parity(vals, ref, what=str(ops))
            assert bi == int(np.argmax(ref)) or near_tie(ref, bi, int(np.argmax(ref)))
    finally:
        ctx.gp_count_set(1)


def test_argmax_first_occurrence_and_edges(ctx):
    g, lo, up = _model(64, 2, "rbf")
    upload_oracle_gp(ctx, 0, g)
    ctx.program_set([[_lib.OP_EI, 0, 0, 0]], [-0.3])
    Xc = _cands(5000, 2, lo, up)
    vals, _, _, bi = ctx.eval(Xc)
    Xd = Xc.copy()
    Xd[4000] = Xd[bi]                      # exact duplicates of the winner, before and after it
    Xd[10] = Xd[bi]
    _, _, bv2, bi2 = ctx.eval(Xd)
    assert bi2 == min(10, bi) and bv2 == vals[bi, 0]
    # N = 0: nothing is touched, index -1
    v0, _, _, b0 = ctx.eval(np.empty((0, 2)))
    assert v0.shape == (0, 1) and b0 == -1
    # determinism / sharding property: one call == two halves, bit for bit
    a, _, _, _ = ctx.eval(Xc[:2500])
    b, _, _, _ = ctx.eval(Xc[2500:])
    assert np.array_equal(np.vstack((a, b)), vals)
    # row independence: a permutation of the candidates permutes the scores
    perm = np.random.RandomState(0).permutation(5000)
    vp, _, _, _ = ctx.eval(Xc[perm])
    assert np.array_equal(vp, vals[perm])


def test_api_errors(ctx):
    g, lo, up = _model(30, 2, "rbf")
    upload_oracle_gp(ctx, 0, g)
    ctx.program_set([[_lib.OP_EI, 0, 0, 0]], [0.0])
    with pytest.raises(_lib.GpfoError):                    # wrong candidate dimension
        ctx.eval(np.zeros((4, 3)))
    with pytest.raises(_lib.GpfoError):                    # program that leaves two values on the stack
        ctx.program_set([[_lib.OP_EI, 0, 0, 0], [_lib.OP_EI, 0, 0, 0]], [0.0])
    with pytest.raises(_lib.GpfoError):                    # slot no GP produces
        ctx.program_set([[_lib.OP_EI, 5, 0, 0]], [0.0])
        ctx.eval(np.zeros((4, 2)))
    ctx.program_set([[_lib.OP_EI, 0, 0, 0]], [0.0])
    with pytest.raises(_lib.GpfoError):                    # unsupported covariance family
        ctx.gp_set(1, g.Xs, g.Ys, 7, [1.0, 1.0], 1.0, 0.1, slot0=1)


@pytest.mark.parametrize("R", [2, 3])
def test_hvpoi_parity(ctx, R):
    """HVPoI over the Pareto cell decomposition (hvpoi.py:98-140): R GPs, front of the predicted means."""
    D, M, N = 4, 80, 1500
    X = synthetic.candidates(M, D, seed=5)
    F = synthetic.dtlz2_objectives(X, R)
    h = synthetic.hypers(D, 1e-3)
    gps = [oracle_gp(X, F[:, [j]], ogp.MATERN52, h) for j in range(R)]
    for j, g in enumerate(gps):
        upload_oracle_gp(ctx, j, g, slot0=j)
    ctx.gp_count_set(R)
    try:
        Fm = np.hstack([g.predict(X)[0] for g in gps])
        p = opareto.Pareto(np.zeros((1, R)))
        p.update(Fm)
        hv = oacq.HVPoI(gps, p.front, p.lb, p.ub)
        ctx.cells_set(hv.pf_ext(), p.lb, p.ub)
        ctx.program_set([[_lib.OP_HVPOI, 0, R, 0]], [])
        Xc = synthetic.candidates(N, D)
        vals, _, bv, bi = ctx.eval(Xc)
        ref = hv.value(Xc)
        assert p.lb.shape[0] > 10 and np.max(ref) > 0
        assert_parity(vals, ref, what="HVPoI R=%d cells=%d" % (R, p.lb.shape[0]))
        assert bi == int(np.argmax(ref)) or near_tie(ref, bi, int(np.argmax(ref)))
        # HVPoI x PoF (constrained multi-objective) through the same program machinery
        ctx.program_set([[_lib.OP_HVPOI, 0, R, 0], [_lib.OP_POF, 0, 0, 0], [_lib.OP_PROD, 2, 0, 0]], [0.8])
        vals2, _, _, _ = ctx.eval(Xc)
        ref2 = ref * oacq.Leaf('pof', gps[0], 0.8).value(Xc)
        assert_parity(vals2, ref2, what="HVPoI*PoF")
    finally:
        ctx.gp_count_set(1)


def test_hvpoi_reference_fixture(ctx):
    """reference testing/unit/test_implementations.py:184-187 + testing/data/vlmop.npz, tolerance decimal=2 as there."""
    import os
    from helpers import GOLDEN, fit_mle
    d = np.load(os.path.join(GOLDEN, "vlmop.npz"))
    gps = [oracle_gp(d['X'], d['Y'][:, [r]], ogp.MATERN32, fit_mle(d['X'], d['Y'][:, [r]], ogp.MATERN32)) for r in range(2)]
    for j, g in enumerate(gps):
        upload_oracle_gp(ctx, j, g, slot0=j)
    ctx.gp_count_set(2)
    try:
        means = np.hstack([ctx.predict(j, d['X'], 1)[0] for j in range(2)])
        p = opareto.Pareto(np.zeros((1, 2)))
        p.update(means)
        hv = oacq.HVPoI(gps, p.front, p.lb, p.ub)
        ctx.cells_set(hv.pf_ext(), p.lb, p.ub)
        ctx.program_set([[_lib.OP_HVPOI, 0, 2, 0]], [])
        vals, _, _, _ = ctx.eval(d['candidates'])
        np.testing.assert_almost_equal(vals, d['scores'], decimal=2)
    finally:
        ctx.gp_count_set(1)


@pytest.mark.parametrize("variant", [20, 21, 19, 12, 11, 4] + EXPERIMENT_VARIANTS)
def test_multi_chunk_variants_agree_with_oracle(ctx, variant):
    """M = 1100 spans several row chunks in every layout (256/512/1024-row chunks, partial last chunk): the resident-panel
    variants (20, 21), the scratch variants (11, 12) and the regenerating variant (4) all meet the oracle."""
    g, lo, up = _model(1100, 6, "matern52", scaled=True)
    upload_oracle_gp(ctx, 0, g)
    fmin = float(np.min(g.predict(g.input_transform.backward(g.Xs))[0]))
    ctx.program_set([[_lib.OP_EI, 0, 0, 0]], [fmin])
    _set_variant_or_skip(ctx, variant)
    try:
        for N in (1, 65, 3000):
            Xc = _cands(N, 6, lo, up)
            vals, _, bv, bi = ctx.eval(Xc)
            ref = oacq.Leaf('ei', g, fmin).value(Xc)
            assert_parity(vals, ref, what="variant %d N=%d" % (variant, N))
            assert bi == int(np.argmax(ref)) or near_tie(ref, bi, int(np.argmax(ref)))
    finally:
        ctx.set_variant(0)


@pytest.mark.parametrize("variant", [20, 21, 19])
@pytest.mark.parametrize("M,D,kind", [(2300, 16, "matern52"), (1541, 3, "rbf"), (3100, 8, "matern32")])
def test_resident_panel_kernel_streaming_groups(ctx, variant, M, D, kind):
    """v5 (csrc/contract5_kernel.cuh) beyond its resident panel: M > 1024 (variant 20) / > 512 (variant 21) regenerates the
    columns right of the panel through the mbarrier-guarded streaming ring, in up to 7 row chunks with a ragged last one;
    D = 16 shrinks the ring to three buffers.  Values, moments (predict) and the fused argmax against the oracle; several
    tiles per CTA (N > 16 * SMs) exercise the phase bookkeeping across tiles."""
    g, lo, up = _model(M, D, kind, scaled=(D == 16))
    upload_oracle_gp(ctx, 0, g)
    fmin = float(np.min(g.predict(g.input_transform.backward(g.Xs))[0]))
    ctx.program_set([[_lib.OP_EI, 0, 0, 0]], [fmin])
    ctx.set_variant(variant)
    try:
        for N in (5, 6000):
            Xc = _cands(N, D, lo, up)
            vals, _, bv, bi = ctx.eval(Xc)
            ref = oacq.Leaf('ei', g, fmin).value(Xc)
            assert_parity(vals, ref, what="variant %d M=%d N=%d" % (variant, M, N))
            assert bv == vals[bi, 0] and bi == int(np.argmax(vals))
            assert bi == int(np.argmax(ref)) or near_tie(ref, bi, int(np.argmax(ref)))
        Xc = _cands(700, D, lo, up, seed=5)
        mean, var = ctx.predict(0, Xc, 1)
        mo, vo = g.predict(Xc)
        assert_parity(mean, mo, what="mean")
        assert_parity(var, vo, scale=g.variance / np.diag(g.output_transform.A)[0] ** 2, what="var")
        a, _, _, _ = ctx.eval(Xc)                       # bitwise repeatable (fixed reduction order, no atomics)
        b, _, _, _ = ctx.eval(Xc)
        assert np.array_equal(a, b)
    finally:
        ctx.set_variant(0)


def test_headline_shape_subset_and_properties(ctx):
    """BASELINE configs[1] at full M (2048, D=8, Matern52): oracle parity on a 20 000-row subset, plus
    size-independent properties on a larger batch (argmax == argmax of the returned scores, halves == whole)."""
    g, lo, up = _model(2048, 8, "matern52")
    upload_oracle_gp(ctx, 0, g)
    fmin = float(np.min(g.predict(g.Xs)[0]))
    ctx.program_set([[_lib.OP_EI, 0, 0, 0]], [fmin])
    N = 200000
    Xc = synthetic.candidates(N, 8)
    vals, _, bv, bi = ctx.eval(Xc)
    sub = np.random.RandomState(9).choice(N, 20000, replace=False)
    sub[0] = bi
    ref = oacq.Leaf('ei', g, fmin).value(Xc[sub])
    assert_parity(vals[sub], ref, what="EI M=2048 subset")
    assert bi == int(np.argmax(vals)) and bv == vals[bi, 0]
    assert int(np.argmax(ref)) == 0 or near_tie(ref, 0, int(np.argmax(ref)))
    a, _, _, _ = ctx.eval(Xc[:N // 2 + 7])
    b, _, _, _ = ctx.eval(Xc[N // 2 + 7:])
    assert np.array_equal(np.vstack((a, b)), vals)
    # the stress conditioning of SURVEY 8d (noise 1e-6, cond ~ 4e7) stays inside the same tolerance
    g2, _, _ = _model(1024, 8, "rbf", noise=1e-6)
    upload_oracle_gp(ctx, 0, g2)
    Xs = synthetic.candidates(3000, 8, seed=8)
    mean, var = ctx.predict(0, Xs, 1)
    mo, vo = g2.predict(Xs)
    from helpers import scale_rel_err
    e_mean, e_var = scale_rel_err(mean, mo), scale_rel_err(var, vo, g2.variance)
    print("cond-4e7 stress (RBF, M=1024, noise 1e-6): scale-relative error mean %.3e, variance %.3e" % (e_mean, e_var))
    # measured on B200 ("stress conditioning"): the bar stays 1e-9
    assert e_mean <= RTOL and e_var <= RTOL, (e_mean, e_var)


# ------------------------------------------------------------------------------------------------------------------
# gradients: d acq / d Xcand against the oracle's closed form (itself checked against torch autograd of
# the reference's op graph in tests/test_oracle_gp.py).  Same scale-relative 1e-9 tolerance, scale = max |grad|.
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["rbf", "matern52", "matern32"])
@pytest.mark.parametrize("leaf,param", [("ei", -0.5), ("poi", -0.5), ("pof", 0.15), ("lcb", 2.0)])
def test_leaf_gradients(ctx, kind, leaf, param):
    g, lo, up = _model(260, 5, kind, scaled=True)
    upload_oracle_gp(ctx, 0, g)
    if leaf in ("ei", "poi"):
        param = float(np.min(g.predict(g.input_transform.backward(g.Xs))[0])) + 0.1
    op = {"ei": _lib.OP_EI, "poi": _lib.OP_POI, "pof": _lib.OP_POF, "lcb": _lib.OP_LCB}[leaf]
    ctx.program_set([[op, 0, 0, 0]], [param])
    for N in (1, 33, 700):
        Xc = _cands(N, 5, lo, up)
        vals, grads, bv, bi = ctx.eval(Xc, want_grads=True)
        rv, rg = oacq.Leaf(leaf, g, param).value_and_grad(Xc)
        assert_parity(vals, rv, what="%s value (grad call) N=%d" % (leaf, N))
        assert_parity(grads, rg, what="%s grad N=%d" % (leaf, N))
        assert grads.shape == (N, 5)
        v_only, _, _, _ = ctx.eval(Xc)
        np.testing.assert_array_equal(v_only, vals)     # reference test_implementations.py:33: evaluate == ewg[0]
        assert bi == int(np.argmax(vals))


def test_product_sum_gradients_multi_gp(ctx):
    D, N = 3, 500
    g1, lo, up = _model(120, D, "matern52", scaled=True)
    g2, _, _ = _model(90, D, "rbf", seed=77, scaled=True)
    for i, g in enumerate((g1, g2)):
        upload_oracle_gp(ctx, i, g, slot0=i)
    ctx.gp_count_set(2)
    try:
        Xc = _cands(N, D, lo, up)
        f1 = float(np.min(g1.predict(g1.input_transform.backward(g1.Xs))[0]))
        ei1, pof2, lcb2 = oacq.Leaf('ei', g1, f1), oacq.Leaf('pof', g2, 2.5), oacq.Leaf('lcb', g2, 1.5)
        cases = [
            ([[_lib.OP_EI, 0, 0, 0], [_lib.OP_POF, 1, 0, 1], [_lib.OP_PROD, 2, 0, 0]], [f1, 2.5], oacq.Agg('prod', [ei1, pof2])),
            ([[_lib.OP_EI, 0, 0, 0], [_lib.OP_POF, 1, 0, 1], [_lib.OP_LCB, 1, 0, 2], [_lib.OP_SUM, 2, 0, 0],
              [_lib.OP_PROD, 2, 0, 0], [_lib.OP_SCALE, 0, 0, 3]], [f1, 2.5, 1.5, 0.25],
             oacq.Agg('prod', [ei1, oacq.Agg('sum', [pof2, lcb2])], scale=0.25)),
        ]
        for ops, params, node in cases:
            ctx.program_set(ops, params)
            vals, grads, _, _ = ctx.eval(Xc, want_grads=True)
            rv, rg = node.value_and_grad(Xc)
            assert_parity(vals, rv, what="value " + str(ops))
            assert_parity(grads, rg, what="grad " + str(ops))
    finally:
        ctx.gp_count_set(1)


@pytest.mark.parametrize("R", [2, 3])
def test_hvpoi_gradients(ctx, R):
    D, M, N = 4, 70, 400
    X = synthetic.candidates(M, D, seed=5)
    F = synthetic.dtlz2_objectives(X, R)
    h = synthetic.hypers(D, 1e-3)
    gps = [oracle_gp(X, F[:, [j]], ogp.MATERN52, h) for j in range(R)]
    for j, g in enumerate(gps):
        upload_oracle_gp(ctx, j, g, slot0=j)
    ctx.gp_count_set(R)
    try:
        p = opareto.Pareto(np.zeros((1, R)))
        p.update(np.hstack([g.predict(X)[0] for g in gps]))
        hv = oacq.HVPoI(gps, p.front, p.lb, p.ub)
        ctx.cells_set(hv.pf_ext(), p.lb, p.ub)
        ctx.program_set([[_lib.OP_HVPOI, 0, R, 0]], [])
        Xc = synthetic.candidates(N, D)
        vals, grads, _, _ = ctx.eval(Xc, want_grads=True)
        rv, rg = hv.value_and_grad(Xc)
        assert_parity(vals, rv, what="HVPoI value R=%d" % R)
        assert_parity(grads, rg, what="HVPoI grad R=%d" % R)
    finally:
        ctx.gp_count_set(1)


def test_gradient_large_m_and_finite_difference(ctx):
    """M = 1100 (three 512-row chunks incl. a partial one) and an independent central finite-difference check."""
    g, lo, up = _model(1100, 8, "matern52")
    upload_oracle_gp(ctx, 0, g)
    fmin = float(np.min(g.predict(g.Xs)[0]))
    ctx.program_set([[_lib.OP_EI, 0, 0, 0]], [fmin])
    Xc = _cands(300, 8, lo, up)
    vals, grads, _, _ = ctx.eval(Xc, want_grads=True)
    rv, rg = oacq.Leaf('ei', g, fmin).value_and_grad(Xc)
    assert_parity(grads, rg, what="EI grad M=1100")
    eps = 1e-6
    for d in (0, 7):
        e = np.zeros(8)
        e[d] = eps
        fd = (ctx.eval(Xc + e)[0] - ctx.eval(Xc - e)[0])[:, 0] / (2 * eps)
        assert np.max(np.abs(fd - grads[:, d])) <= 1e-5 * np.max(np.abs(grads))


# ------------------------------------------------------------------------------------------------------------------
# shape / parameter edge cases
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("D", [1, 2, 7, 33, 64])
def test_input_dimensions(ctx, D):
    X, Y = synthetic.training_set(90, D)
    g = oracle_gp(X, Y, ogp.MATERN52, synthetic.hypers(D))
    upload_oracle_gp(ctx, 0, g)
    ctx.program_set([[_lib.OP_LCB, 0, 0, 0]], [1.0])
    Xc = synthetic.candidates(500, D)
    vals, grads, _, _ = ctx.eval(Xc, want_grads=True)
    rv, rg = oacq.Leaf('lcb', g, 1.0).value_and_grad(Xc)
    assert_parity(vals, rv, what="LCB D=%d" % D)
    assert_parity(grads, rg, what="LCB grad D=%d" % D)


def test_isotropic_lengthscale_and_kernel_families(ctx):
    """Non-ARD kernels (one lengthscale for all inputs), as create_vlmop2_model builds them (testing/utility.py:61-62)."""
    X, Y = synthetic.training_set(120, 3)
    for kind in (ogp.RBF, ogp.MATERN52, ogp.MATERN32):
        h = dict(lengthscales=np.array([0.8]), variance=0.9, noise_variance=1e-3)
        g = oracle_gp(X, Y, kind, h)
        upload_oracle_gp(ctx, 0, g)
        Xc = synthetic.candidates(300, 3)
        mean, var = ctx.predict(0, Xc, 1)
        mo, vo = g.predict(Xc)
        assert_parity(mean, mo, what="mean kind %d" % kind)
        assert_parity(var, vo, scale=g.variance, what="var kind %d" % kind)


def test_leaf_on_second_output_column_and_slot_offsets(ctx):
    """Programs may reference any moment slot: a 3-output GP in slots 0..2 and a 1-output GP in slot 3."""
    D = 3
    X, Y = synthetic.training_set(80, D, constraint=True)
    Y3 = np.hstack((Y, (Y[:, [0]] * Y[:, [1]])))
    h = synthetic.hypers(D)
    g3 = oracle_gp(X, Y3, ogp.RBF, h)
    g1 = oracle_gp(X, Y[:, [1]], ogp.MATERN32, h)
    upload_oracle_gp(ctx, 0, g3, slot0=0)
    upload_oracle_gp(ctx, 1, g1, slot0=3)
    ctx.gp_count_set(2)
    try:
        # EI on column 2 of the first GP times PoF of the second GP, plus the mean of column 1
        ctx.program_set([[_lib.OP_EI, 2, 0, 0], [_lib.OP_POF, 3, 0, 1], [_lib.OP_PROD, 2, 0, 0], [_lib.OP_MEAN, 1, 0, 0],
                         [_lib.OP_SUM, 2, 0, 0]], [0.1, 0.0])
        Xc = synthetic.candidates(400, D)
        vals, grads, _, _ = ctx.eval(Xc, want_grads=True)
        m3, v3 = g3.predict(Xc)
        ei = oacq.Leaf('ei', g3, 0.1)._value(m3[:, [2]], v3[:, [2]])
        pof = oacq.Leaf('pof', g1, 0.0).value(Xc)
        assert_parity(vals, ei * pof + m3[:, [1]], what="multi-output slots")
        eps = 1e-6
        e = np.zeros(D)
        e[1] = eps
        fd = (ctx.eval(Xc + e)[0] - ctx.eval(Xc - e)[0])[:, 0] / (2 * eps)
        assert np.max(np.abs(fd - grads[:, 1])) <= 2e-5 * np.max(np.abs(grads))
    finally:
        ctx.gp_count_set(1)


def test_large_candidate_batch_device_resident(ctx):
    """10^7 candidates resident on the device (config #3's per-job size), M small: exercises 64-bit indexing and the
    device-pointer path of gpfo_eval; checked through sub-batch consistency."""
    import torch
    g, lo, up = _model(64, 4, "rbf")
    upload_oracle_gp(ctx, 0, g)
    ctx.program_set([[_lib.OP_EI, 0, 0, 0]], [-0.3])
    N = 10_000_000
    gen = torch.Generator(device="cuda").manual_seed(7)
    Xd = torch.rand((N, 4), dtype=torch.float64, device="cuda", generator=gen)
    vd = torch.empty(N, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()                  # the library launches on its own stream: torch's generator must be done
    bv, bi = ctx.eval_device(Xd.data_ptr(), N, 4, vd.data_ptr())
    torch.cuda.synchronize()
    assert int(torch.argmax(vd)) == bi and float(vd[bi]) == bv
    tail = Xd[N - 1000:].cpu().numpy()
    ref = oacq.Leaf('ei', g, -0.3).value(tail)
    assert_parity(vd[N - 1000:].cpu().numpy()[:, None], ref, what="tail of a 1e7 batch")


# ------------------------------------------------------------------------------------------------------------------
# row-split launches (few candidates: the rows of the block stream are spread over the SMs, partial sums added by a
# finish kernel).  Covers every split shape: forward + gradient small tiles (N <= 592), wide gradient tiles
# (592 < N <= 2368), partial last chunk (M % 128 != 0), multi-output, and agreement with the un-split kernel.
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,D", [(300, 3), (1100, 8), (2048, 8)])
def test_row_split_small_batches(ctx, M, D):
    g, lo, up = _model(M, D, "matern52", scaled=True)
    upload_oracle_gp(ctx, 0, g)
    fmin = float(np.min(g.predict(g.input_transform.backward(g.Xs))[0])) + 0.05
    ctx.program_set([[_lib.OP_EI, 0, 0, 0]], [fmin])
    leaf = oacq.Leaf('ei', g, fmin)
    for N in (1, 7, 8, 9, 100, 592, 593, 1500):
        Xc = _cands(N, D, lo, up, seed=N)
        vals, grads, bv, bi = ctx.eval(Xc, want_grads=True)
        rv, rg = leaf.value_and_grad(Xc)
        assert_parity(vals, rv, what="split value M=%d N=%d" % (M, N))
        assert_parity(grads, rg, what="split grad M=%d N=%d" % (M, N))
        assert bi == int(np.argmax(vals)) and bv == vals[bi, 0]
        mean, var = ctx.predict(0, Xc, 1)
        rm, rvv = g.predict(Xc)
        assert_parity(mean, rm, what="split predict mean")
        assert_parity(var, rvv, scale=g.variance / g.output_transform.A[0, 0] ** 2, what="split predict var")
        ctx.set_variant(6)                    # explicit variant: same inputs through the un-split kernels
        try:
            v6, g6, _, i6 = ctx.eval(Xc, want_grads=True)
        finally:
            ctx.set_variant(0)
        # different summation orders (the number of row splits follows the SM count of the box): well inside the 1e-9 bar
        assert_parity(vals, v6, tol=1e-11, what="split vs un-split value")
        assert_parity(grads, g6, tol=1e-10, what="split vs un-split grad")


# ------------------------------------------------------------------------------------------------------------------
# latency path (a handful of candidates: 16-row contraction kernels on two streams + one finish kernel; row-split DMMA
# kernels + the same finish kernel up to N = 592): every candidate-count class (NV = 1 / 4 / 8, several tiles, the hand-over
# to the row-split kernels at N = 64 / 65), two GPs with different kernels and M, a multi-output GP, repeated calls (the third
# call of a shape replays a CUDA graph with a forked second stream) and evaluate == evaluate_with_gradients[0] bitwise.
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M1,M2,D", [(50, 37, 3), (300, 120, 5), (1100, 515, 8)])
def test_latency_path_small_calls(ctx, M1, M2, D):
    g1, lo, up = _model(M1, D, "matern52", scaled=True)
    g2, _, _ = _model(M2, D, "rbf", seed=77, scaled=True)
    for i, g in enumerate((g1, g2)):
        upload_oracle_gp(ctx, i, g, slot0=i)
    ctx.gp_count_set(2)
    try:
        f1 = float(np.min(g1.predict(g1.input_transform.backward(g1.Xs))[0])) + 0.05
        ei1, pof2, lcb2 = oacq.Leaf('ei', g1, f1), oacq.Leaf('pof', g2, 2.5), oacq.Leaf('lcb', g2, 1.5)
        ctx.program_set([[_lib.OP_EI, 0, 0, 0], [_lib.OP_POF, 1, 0, 1], [_lib.OP_LCB, 1, 0, 2], [_lib.OP_SUM, 2, 0, 0],
                         [_lib.OP_PROD, 2, 0, 0], [_lib.OP_SCALE, 0, 0, 3]], [f1, 2.5, 1.5, 0.25])
        node = oacq.Agg('prod', [ei1, oacq.Agg('sum', [pof2, lcb2])], scale=0.25)
        for N in (1, 2, 4, 5, 8, 9, 31, 64, 65, 130):
            Xc = _cands(N, D, lo, up, seed=100 + N)
            rv, rg = node.value_and_grad(Xc)
            for rep in range(3):                                  # plain, capture, replay
                vals, grads, bv, bi = ctx.eval(Xc, want_grads=True)
                assert_parity(vals, rv, what="latency path value N=%d rep %d" % (N, rep))
                assert_parity(grads, rg, what="latency path grad N=%d rep %d" % (N, rep))
                assert bi == int(np.argmax(vals)) and bv == vals[bi, 0]
                v_only, _, bv2, bi2 = ctx.eval(Xc)
                np.testing.assert_array_equal(v_only, vals)
                assert (bv2, bi2) == (bv, bi)
    finally:
        ctx.gp_count_set(1)


def test_latency_path_multi_output_slots(ctx):
    """The latency path on a 3-output GP plus a 1-output GP: same rows as a 400-row call through the standard kernels."""
    D = 3
    X, Y = synthetic.training_set(80, D, constraint=True)
    Y3 = np.hstack((Y, (Y[:, [0]] * Y[:, [1]])))
    h = synthetic.hypers(D)
    upload_oracle_gp(ctx, 0, oracle_gp(X, Y3, ogp.RBF, h), slot0=0)
    upload_oracle_gp(ctx, 1, oracle_gp(X, Y[:, [1]], ogp.MATERN32, h), slot0=3)
    ctx.gp_count_set(2)
    try:
        ctx.program_set([[_lib.OP_EI, 2, 0, 0], [_lib.OP_POF, 3, 0, 1], [_lib.OP_PROD, 2, 0, 0], [_lib.OP_MEAN, 1, 0, 0],
                         [_lib.OP_SUM, 2, 0, 0]], [0.1, 0.0])
        Xc = synthetic.candidates(700, D)
        vals, grads, _, _ = ctx.eval(Xc, want_grads=True)        # 700 rows: beyond every small-call path
        for N in (1, 3, 8, 20, 64, 100):
            v, g, bv, bi = ctx.eval(Xc[:N], want_grads=True)
            assert_parity(v, vals[:N], tol=1e-11, what="latency path vs standard kernels, value N=%d" % N)
            assert_parity(g, grads[:N], tol=1e-10, what="latency path vs standard kernels, grad N=%d" % N)
            assert bi == int(np.argmax(v)) and bv == v[bi, 0]
    finally:
        ctx.gp_count_set(1)


# ------------------------------------------------------------------------------------------------------------------
# stored-A gradient pass (large batches, N >= 64 * 4 * SMs): forward launches park A = L^-1 Kx, backward launches
# contract it with the reversed transposed stream.  Checked against the oracle on a row subset, against the dense
# K^-1 pass on every row, across several slabs, for a two-GP product and for HVPoI.
# ------------------------------------------------------------------------------------------------------------------
def _with_env(name, value, fn):
    import os
    old = os.environ.get(name)
    os.environ[name] = value
    try:
        return fn()
    finally:
        if old is None:
            del os.environ[name]
        else:
            os.environ[name] = old


@pytest.mark.parametrize("M,D,kind", [(1100, 8, "matern52"), (300, 3, "rbf"), (515, 32, "matern32")])
def test_stored_a_gradients_large_batch(ctx, M, D, kind):
    N = 40000
    g, lo, up = _model(M, D, kind, scaled=True)
    upload_oracle_gp(ctx, 0, g)
    fmin = float(np.min(g.predict(g.input_transform.backward(g.Xs))[0])) + 0.05
    ctx.program_set([[_lib.OP_EI, 0, 0, 0]], [fmin])
    Xc = _cands(N, D, lo, up)
    ctx.set_variant(12)            # the large-batch kernel family whatever the SM count of the box (automatic at N >= 256 SMs)
    try:
        _stored_a_checks(ctx, g, fmin, Xc, N)
    finally:
        ctx.set_variant(0)


def _stored_a_checks(ctx, g, fmin, Xc, N):
    vals, grads, bv, bi = ctx.eval(Xc, want_grads=True)
    assert ctx.timings()["launches"] == 4                         # forward + partials + backward + argmax: one slab
    sub = np.arange(0, N, 97)
    rv, rg = oacq.Leaf('ei', g, fmin).value_and_grad(Xc[sub])
    assert_parity(vals[sub], rv, what="stored-A value")
    assert_parity(grads[sub], rg, what="stored-A grad")
    v_only, _, bv2, bi2 = ctx.eval(Xc)
    np.testing.assert_array_equal(v_only, vals)
    assert bi == bi2 == int(np.argmax(vals)) and bv == bv2
    vd, gd, _, _ = _with_env("GPFO_GRAD_DENSE", "1", lambda: ctx.eval(Xc, want_grads=True))
    np.testing.assert_array_equal(vd, vals)
    assert_parity(grads, gd, tol=1e-10, what="stored-A vs dense K^-1 pass")
    # several slabs (bounded A scratch): same numbers, slab by slab
    vs, gs, bvs, bis = _with_env("GPFO_ASTORE_BUDGET_MB", "1", lambda: ctx.eval(Xc, want_grads=True))
    assert ctx.timings()["launches"] > 4
    np.testing.assert_array_equal(vs, vals)
    np.testing.assert_array_equal(gs, grads)
    assert bis == bi and bvs == bv


def test_stored_a_product_of_two_gps_and_hvpoi(ctx):
    D, N = 4, 38400
    g1, lo, up = _model(200, D, "matern52", scaled=True)
    g2, _, _ = _model(333, D, "rbf", seed=77, scaled=True, R=1)
    for i, g in enumerate((g1, g2)):
        upload_oracle_gp(ctx, i, g, slot0=i)
    ctx.gp_count_set(2)
    try:
        Xc = _cands(N, D, lo, up)
        sub = np.arange(5, N, 101)
        ctx.set_variant(12)
        f1 = float(np.min(g1.predict(g1.input_transform.backward(g1.Xs))[0]))
        ctx.program_set([[_lib.OP_EI, 0, 0, 0], [_lib.OP_POF, 1, 0, 1], [_lib.OP_PROD, 2, 0, 0]], [f1, 2.5])
        vals, grads, _, bi = _with_env("GPFO_ASTORE_BUDGET_MB", "8", lambda: ctx.eval(Xc, want_grads=True))
        node = oacq.Agg('prod', [oacq.Leaf('ei', g1, f1), oacq.Leaf('pof', g2, 2.5)])
        rv, rg = node.value_and_grad(Xc[sub])
        assert_parity(vals[sub], rv, what="stored-A product value")
        assert_parity(grads[sub], rg, what="stored-A product grad")
        assert bi == int(np.argmax(vals))
        # HVPoI over the two means as objectives
        X = g1.input_transform.backward(g1.Xs)
        p = opareto.Pareto(np.zeros((1, 2)))
        gps = [g1, g2]
        p.update(np.hstack([g.predict(X)[0] for g in gps]))
        hv = oacq.HVPoI(gps, p.front, p.lb, p.ub)
        ctx.cells_set(hv.pf_ext(), p.lb, p.ub)
        ctx.program_set([[_lib.OP_HVPOI, 0, 2, 0]], [])
        vals, grads, _, bi = ctx.eval(Xc, want_grads=True)
        rv, rg = hv.value_and_grad(Xc[sub])
        assert_parity(vals[sub], rv, what="stored-A HVPoI value")
        assert_parity(grads[sub], rg, what="stored-A HVPoI grad")
        assert bi == int(np.argmax(vals))
    finally:
        ctx.set_variant(0)
        ctx.gp_count_set(1)


# ------------------------------------------------------------------------------------------------------------------
# Min-value entropy search leaf (reference mes.py:96-102): values and gradients, all log_ndtr segments
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["rbf", "matern52"])
def test_mes_leaf_values_and_gradients(ctx, kind):
    g, lo, up = _model(260, 5, kind, scaled=True)
    upload_oracle_gp(ctx, 0, g)
    fmin = float(np.min(g.predict(g.input_transform.backward(g.Xs))[0]))
    m0, v0 = g.predict(_cands(700, 5, lo, up))
    # samples below the best mean (the usual case), one far below (gamma > 8) and one above the means by ~25 predictive
    # standard deviations (gamma < -20: series segment; where cdf underflows both sides give NaN, as the reference would)
    samples = np.concatenate((fmin - np.array([0.02, 0.1, 0.3, 0.8, 2.0, 60.0, 0.5, 1.0]),
                              [np.median(m0) + 25.0 * np.median(np.sqrt(v0))]))
    node = oacq.MES(g, samples)
    for nsamp in (8, 9):
        ctx.program_set([[_lib.OP_MES, 0, nsamp, 0]], list(samples))
        node = oacq.MES(g, samples[:nsamp])
        Xtr = g.input_transform.backward(g.Xs)
        xbest = Xtr[[int(np.argmin(g.predict(Xtr)[0]))]]
        for N in (1, 33, 700, 5000):
            # N = 1: a point next to the incumbent (gamma ~ 1; far from the data MES is ~1e-60 and its own scale)
            Xc = _cands(N, 5, lo, up) if N > 1 else np.clip(xbest + 0.05, lo, up)
            vals, grads, bv, bi = ctx.eval(Xc, want_grads=True)
            with np.errstate(all='ignore'):
                rv, rg = node.value_and_grad(Xc)
            # compare where both sides are finite (cdf underflow makes 0/0 on either side, a few ulp apart in gamma)
            ok = np.isfinite(rv[:, 0]) & np.isfinite(vals[:, 0]) & np.all(np.isfinite(rg), axis=1) & np.all(np.isfinite(grads), axis=1)
            pm, pv = g.predict(Xc)                         # ... and where cdf(gamma) is a normal number (gamma > -37)
            gmin = np.min((pm - node.samples[None, :]) / np.sqrt(pv), axis=1)
            ok &= gmin > -37.0
            # d/dgamma of gamma pdf / (2 cdf) - log cdf cancels from O(gamma^3) terms down to O(1/gamma) for gamma << 0: an
            # erfc difference of 1e-13 between two libms is amplified by gamma^4 / 2, so the 1e-9 gradient bar is meaningful
            # only for gamma > -10 (always the case for y* samples below the posterior means); beyond that 1e-4.
            well = ok & (gmin > -10.0)
            assert (N == 1 or ok.mean() > 0.5) and (nsamp == 9 or ok.all())
            v_only, _, _, bi2 = ctx.eval(Xc)
            np.testing.assert_array_equal(v_only, vals)
            if ok.any():
                assert_parity(vals[ok], rv[ok], what="MES value N=%d" % N)
                if well.any():
                    assert_parity(grads[well], rg[well], what="MES grad N=%d" % N)
                assert_parity(grads[ok], rg[ok], tol=1e-4, what="MES grad (ill-conditioned rows) N=%d" % N)
                assert bi == bi2 == int(np.nanargmax(vals))
            else:
                assert bi == bi2 == -1                      # NaN never wins (reference: np.argmin over NaN is undefined)
    gam = (m0 - samples[8]) / np.sqrt(v0)
    assert np.any((gam < -20.0) & (gam > -37.0)) and np.any((m0 - samples[5]) / np.sqrt(v0) > 8.0)     # segments were hit
    with pytest.raises(_lib.GpfoError):
        ctx.program_set([[_lib.OP_MES, 0, 5, 2]], [0.0] * 6)       # sample range runs past params


# ------------------------------------------------------------------------------------------------------------------
# BASELINE configs at their full per-GPU sizes: oracle parity on a row subset + size-independent properties
# (a slice scored alone reproduces its rows of the whole batch bit for bit; arg-max == arg-max of the returned scores;
# the fused arg-max call without scores returns the same record)
# ------------------------------------------------------------------------------------------------------------------
def test_config3_full_size_constrained_ei(ctx):
    """configs[2]: EI x PoF, 2 GPs, M = 4096, D = 16, one GPU's 1/8 share of the 1e7 candidates."""
    M, D, N = 4096, 16, 1250000
    X, Y2 = synthetic.training_set(M, D, constraint=True)
    h = synthetic.hypers(D)
    g1 = oracle_gp(X, Y2[:, [0]], ogp.MATERN52, h)
    g2 = oracle_gp(X, Y2[:, [1]], ogp.MATERN52, h)
    for i, g in enumerate((g1, g2)):
        upload_oracle_gp(ctx, i, g, slot0=i)
    ctx.gp_count_set(2)
    try:
        ctx.program_set([[_lib.OP_EI, 0, 0, 0], [_lib.OP_POF, 1, 0, 1], [_lib.OP_PROD, 2, 0, 0]], [-1.0, 0.0])
        Xc = synthetic.candidates(N, D)
        vals, _, bv, bi = ctx.eval(Xc)
        sub = np.random.RandomState(2).choice(N, 400, replace=False)
        sub[0] = bi
        ref = oacq.Agg('prod', [oacq.Leaf('ei', g1, -1.0), oacq.Leaf('pof', g2, 0.0)]).value(Xc[sub])
        assert_parity(vals[sub], ref, what="config 3 subset")
        assert bi == int(np.argmax(vals)) and bv == vals[bi, 0]
        _, _, bv2, bi2 = ctx.eval(Xc, want_values=False)
        assert (bv2, bi2) == (bv, bi)
        lo_, hi_ = 777777, 777777 + 9472                      # a slice scored alone (different tile width, same numbers)
        part, _, _, _ = ctx.eval(Xc[lo_:hi_])
        assert_parity(part, vals[lo_:hi_], tol=1e-11, what="config 3 slice vs whole")
    finally:
        ctx.gp_count_set(1)


def test_config4_full_size_hvpoi(ctx):
    """configs[3]: HVPoI, 3 objectives, M = 1024 per GP, 1e6 candidates, Pareto front of ~200 points."""
    from gpflowopt_b200 import pareto as gpareto
    M, D, R, N = 1024, 6, 3, 1000000                       # D = 6 as SURVEY 8(d) states
    X = synthetic.candidates(M, D, seed=5)
    F = synthetic.dtlz2_objectives(X, R, g_scale=0.4)
    h = synthetic.hypers(D, 1e-3)
    gps = [oracle_gp(X, F[:, [j]], ogp.MATERN52, h) for j in range(R)]
    for j, g in enumerate(gps):
        upload_oracle_gp(ctx, j, g, slot0=j)
    ctx.gp_count_set(R)
    try:
        Fm = np.hstack([ctx.predict(j, X, 1)[0] for j in range(R)])
        p = gpareto.Pareto(np.zeros((1, R)))
        p.update(Fm)
        hv = oacq.HVPoI(gps, np.asarray(p.front), np.asarray(p.bounds.lb), np.asarray(p.bounds.ub))
        assert 150 <= p.front.shape[0] <= 260
        mlb, mub = gpareto.merge_cells(hv.pf_ext(), hv.lb, hv.ub)          # what HVProbabilityOfImprovement uploads
        ctx.cells_set(hv.pf_ext(), mlb, mub)
        ctx.program_set([[_lib.OP_HVPOI, 0, R, 0]], [])
        Xc = synthetic.candidates(N, D)
        vals, _, bv, bi = ctx.eval(Xc)
        sub = np.random.RandomState(3).choice(N, 200, replace=False)
        sub[0] = bi
        assert_parity(vals[sub], hv.value(Xc[sub]), what="config 4 subset (reference cell tables on the oracle side)")
        assert bi == int(np.argmax(vals)) and bv == vals[bi, 0]
        part, _, _, _ = ctx.eval(Xc[123456:123456 + 40000])
        assert_parity(part, vals[123456:123456 + 40000], tol=1e-11, what="config 4 slice vs whole")
    finally:
        ctx.gp_count_set(1)


# ------------------------------------------------------------------------------------------------------------------
# BASELINE configs[0] as stated and configs[4] as a test grid
# ------------------------------------------------------------------------------------------------------------------
def test_config1_branin_mc1000_as_stated(ctx):
    """configs[0]: ExpectedImprovement on 2-D Branin ([-5, 10] x [0, 15]), GPR RBF-ARD with 20 training points, seeded
    MCOptimizer(domain, 1000) (gpflowopt/optim.py:152-164).  Same np.random stream on both sides: the selected ROW must be the
    one oracle.acq.mc_optimize picks (first-occurrence argmin of -EI) and `fun` must agree to 1e-9."""
    import gpflowopt_b200 as gpfo
    from gpflowopt_b200.acquisition import ExpectedImprovement
    from gpflowopt_b200.bo import _InverseAcquisition
    from gpflowopt_b200.design import RandomDesign
    from gpflowopt_b200.domain import ContinuousParameter
    from gpflowopt_b200.optim import MCOptimizer

    def branin(x):
        x1, x2 = x[:, 0], x[:, 1]
        a, b, c, r, s, t = 1, 5.1 / (4 * np.pi ** 2), 5 / np.pi, 6, 10, 1 / (8 * np.pi)
        return (a * (x2 - b * x1 ** 2 + c * x1 - r) ** 2 + s * (1 - t) * np.cos(x1) + s)[:, None]

    domain = ContinuousParameter('x1', -5, 10) + ContinuousParameter('x2', 0, 15)
    rs = np.random.RandomState(7)
    X = domain.lower + rs.rand(20, 2) * (domain.upper - domain.lower)
    Y = branin(X)
    hyp = dict(lengthscales=np.array([0.3, 0.4]), variance=1.0, noise_variance=1e-3)
    for scaled in (True, False):                    # BayesianOptimizer's default (domain -> unit cube, normalised Y) and raw
        ell = hyp["lengthscales"] if scaled else hyp["lengthscales"] * (domain.upper - domain.lower)
        var = hyp["variance"] if scaled else float(np.var(Y))
        model = gpfo.GPR(X, Y, gpfo.RBF(2, variance=var, lengthscales=ell, ARD=True), noise_variance=hyp["noise_variance"])
        acq = ExpectedImprovement(model)
        acq.optimize_restarts = 0
        if scaled:
            acq.enable_scaling(domain)
        g = oracle_gp(X, Y, ogp.RBF, dict(lengthscales=ell, variance=var, noise_variance=hyp["noise_variance"]),
                      lower=domain.lower if scaled else None, upper=domain.upper if scaled else None, normalize=scaled)
        node = oacq.Leaf('ei', g, oacq.fmin_of(g, X))
        np.random.seed(11)
        res = MCOptimizer(domain, 1000).optimize(_InverseAcquisition(acq, False))
        np.random.seed(11)
        points = RandomDesign(1000, domain).generate()                # the stream MCOptimizer consumed
        x_ref, f_ref, idx_ref, evals = oacq.mc_optimize(node, points)
        assert res.success and res.nfev == 1000 and res.x.shape == (1, 2) and res.fun.shape == (1, 1)
        assert np.array_equal(res.x, x_ref), "selected row differs from the oracle's (scaled=%s)" % scaled
        assert abs(res.fun[0, 0] - f_ref[0, 0]) <= RTOL * np.max(np.abs(evals))
        assert_parity(acq.evaluate(points), -evals, what="config 1 scores, scaled=%s" % scaled)
        acq._get_engine().close()


@pytest.mark.parametrize("M", [512, 1024, 2048, 4096, 8192])
def test_config5_sweep_values_and_gradients(ctx, M):
    """configs[4] at reduced N: every M of the sweep, values and evaluate_with_gradients, through the batch sizes that select
    the large-batch kernels (resident-panel forward, stored-A gradient pass), against an oracle subset."""
    D = 8
    g, lo, up = _model(M, D, "matern52")
    upload_oracle_gp(ctx, 0, g)
    fmin = -1.5
    ctx.program_set([[_lib.OP_EI, 0, 0, 0]], [fmin])
    node = oacq.Leaf('ei', g, fmin)
    N = 40000 if M <= 4096 else 38000                       # >= 256 candidates per SM: the automatic large-batch path
    Xc = _cands(N, D, lo, up, seed=M)
    vals, _, bv, bi = ctx.eval(Xc)
    nsub = 600 if M <= 2048 else 150
    sub = np.random.RandomState(M).choice(N, nsub, replace=False)
    sub[0] = bi
    ref = node.value(Xc[sub])
    assert_parity(vals[sub], ref, what="values M=%d" % M)
    assert bi == int(np.argmax(vals)) and bv == vals[bi, 0]
    vg, grads, _, big = ctx.eval(Xc, want_grads=True)
    assert big == bi
    assert_parity(vg[sub], ref, what="values of the gradient call M=%d" % M)
    _, rg = node.value_and_grad(Xc[sub[:100]])
    assert_parity(grads[sub[:100]], rg, what="gradients M=%d" % M)
