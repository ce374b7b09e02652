"""Test-only helpers: tolerances, oracle/product model factories, a small MLE fit for the vlmop fixture."""
import os

import numpy as np
from scipy.linalg import cho_solve, cholesky
from scipy.optimize import minimize

from oracle import gp as ogp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# BASELINE.json north_star: "within rel. 1e-9 (float64)".  Two exact-arithmetic-equivalent CPU formulations already
# differ by up to 3e-7 pointwise where values cross zero, so relative means relative to the
# batch scale: |d| <= 1e-9 * max|ref|  (variances: relative to the prior variance).
RTOL = 1e-9


def scale_rel_err(got, ref, scale=None):
    ref = np.asarray(ref)
    s = np.max(np.abs(ref)) if scale is None else scale
    return float(np.max(np.abs(np.asarray(got) - ref)) / max(s, 1e-300))


def assert_parity(got, ref, scale=None, tol=RTOL, what=""):
    err = scale_rel_err(got, ref, scale)
    assert err <= tol, "%s: scale-relative error %.3e > %.1e" % (what, err, tol)


def near_tie(ref, idx_a, idx_b, tol=RTOL):
    """Documented near-tie rule for argmax identity: both picks are within tol * scale of each other."""
    ref = np.asarray(ref).ravel()
    return abs(ref[idx_a] - ref[idx_b]) <= tol * np.max(np.abs(ref))


def oracle_gp(X, Y, kind, hyp, lower=None, upper=None, normalize=False):
    return ogp.ScaledGPR(X, Y, kind, hyp["lengthscales"], hyp["variance"], hyp["noise_variance"],
                         domain_lower=lower, domain_upper=upper, normalize_Y=normalize)


def upload_oracle_gp(ctx, gp_id, g, slot0=0):
    """Feeds the CUDA library exactly the state the oracle model holds."""
    in_scale, in_shift = np.diag(g.input_transform.A), g.input_transform.b
    out_scale, out_shift = np.diag(g.output_transform.A), g.output_transform.b
    ctx.gp_set(gp_id, g.Xs, g.Ys, g.kind, g.lengthscales, g.variance, g.noise_variance, in_scale, in_shift,
               out_scale, out_shift, slot0)


def fit_mle(X, Y, kind, ard=False, restarts=3, seed=0):
    """Maximum marginal likelihood fit (log-parameters, L-BFGS-B) -- stands in for GPflow's model.optimize() when
    anchoring the oracle on the reference's vlmop fixture, whose hyper-parameters come from such a fit."""
    M, D = X.shape
    nl = D if ard else 1

    def nlml(theta):
        var, ell, noise = np.exp(theta[0]), np.exp(theta[1:1 + nl]), np.exp(theta[-1])
        K = ogp.kern_K(kind, var, ell, X) + (noise + 1e-10) * np.eye(M)
        try:
            L = cholesky(K, lower=True)
        except np.linalg.LinAlgError:
            return 1e10
        alpha = cho_solve((L, True), Y)
        return float(0.5 * np.sum(Y * alpha) + Y.shape[1] * np.sum(np.log(np.diag(L))) + 0.5 * M * Y.shape[1] * np.log(2 * np.pi))

    rs = np.random.RandomState(seed)
    best = None
    for r in range(restarts):
        x0 = np.zeros(2 + nl) if r == 0 else rs.randn(2 + nl)
        res = minimize(nlml, x0, method="L-BFGS-B", bounds=[(-12, 8)] * (2 + nl))
        if best is None or res.fun < best.fun:
            best = res
    th = best.x
    return dict(variance=float(np.exp(th[0])), lengthscales=np.exp(th[1:1 + nl]), noise_variance=float(np.exp(th[-1])))
