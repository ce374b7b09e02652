"""Oracle: acquisition functions over candidate sets, restated in NumPy float64.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  PARITY UNPINNED at 1e-9 for the Normal cdf/pdf
forms (tf.contrib.distributions.Normal is TF1, absent here); pinned only loosely by
testing/data/vlmop.npz (decimal=2) and the ordinal tests in testing/unit/test_implementations.py.

Pinned above the TF layer: tests/golden/acq_golden.npz holds the results of the reference's own build_acquisition methods
executed on fixed moments (NumPy stand-ins for the TF primitives, tests/golden/gen_acq_golden.py); the compositions below
reproduce them bit for bit (tests/test_oracle_gp.py::test_acquisition_compositions_against_reference_executed_graphs).

Restated reference sites (relative to /root/reference):
  * ExpectedImprovement.build_acquisition      gpflowopt/acquisition/ei.py:70-79  (fmin: ei.py:63-68)
  * ProbabilityOfImprovement.build_acquisition gpflowopt/acquisition/poi.py:63-67
  * ProbabilityOfFeasibility.build_acquisition gpflowopt/acquisition/pof.py:81-85 (feasible: pof.py:63-79)
  * LowerConfidenceBound.build_acquisition     gpflowopt/acquisition/lcb.py:54-57
  * AcquisitionSum / Product / MCMC mean       gpflowopt/acquisition/acquisition.py:360-362, 446-448
  * HVProbabilityOfImprovement                 gpflowopt/acquisition/hvpoi.py:79-140
  * MinValueEntropySearch.build_acquisition    gpflowopt/acquisition/mes.py:96-102 (Gumbel sampling of y*: mes.py:63-94)
  * tf.gradients(acq, Xcand)                   gpflowopt/acquisition/acquisition.py:256-257
  * MCOptimizer._optimize (argmin)             gpflowopt/optim.py:155-164
"""
import numpy as np
from scipy.special import erf, erfc

STABILITY = 1e-6          # gpflow settings.numerics.jitter_level (ei.py:24, poi.py:23, pof.py:23, hvpoi.py:24)
HALF_SQRT_2 = 0.5 * np.sqrt(2.0)
HALF_LOG_2PI = 0.5 * np.log(2.0 * np.pi)


def ndtr(t):
    """TF1 ``_ndtr``: erf branch for |t|/sqrt2 < sqrt(1/2), erfc branches otherwise."""
    t = np.asarray(t, dtype=np.float64)
    w = t * HALF_SQRT_2
    z = np.abs(w)
    with np.errstate(invalid='ignore'):
        y = np.where(z < HALF_SQRT_2, 1.0 + erf(w), np.where(w > 0.0, 2.0 - erfc(z), erfc(z)))
    return 0.5 * y


def normal_cdf(x, loc, scale):
    return ndtr((x - loc) / scale)


def normal_prob(x, loc, scale):
    """TF1 Normal.prob = exp(-0.5 z^2 - (0.5 log 2pi + log scale))."""
    z = (x - loc) / scale
    return np.exp(-0.5 * np.square(z) - (HALF_LOG_2PI + np.log(scale)))


def std_pdf(z):
    return np.exp(-0.5 * np.square(z) - HALF_LOG_2PI)


LOGNDTR_LOWER, LOGNDTR_UPPER = -20.0, 8.0     # TF1 special_math float64 segments (RECALLED)


def _log_ndtr_series(x):
    """TF1 ``_log_ndtr_asymptotic_series`` with series_order = 3: 1 + 3/x^4 - (1/x^2 + 15/x^6)."""
    x2 = np.square(x)
    x4 = x2 * x2
    x6 = x4 * x2
    even_sum = 3.0 / x4
    odd_sum = 1.0 / x2 + 15.0 / x6
    return 1.0 + even_sum - odd_sum


def log_ndtr(x):
    """TF1 ``log_ndtr`` (Normal.log_cdf): -ndtr(-x) above 8, log(ndtr(x)) in (-20, 8], asymptotic series below."""
    x = np.asarray(x, dtype=np.float64)
    xl = np.minimum(x, LOGNDTR_LOWER)
    with np.errstate(divide='ignore', invalid='ignore'):
        lower = (-0.5 * np.square(xl) - np.log(-xl) - HALF_LOG_2PI) + np.log(_log_ndtr_series(xl))
        mid = np.log(ndtr(np.maximum(x, LOGNDTR_LOWER)))
    return np.where(x > LOGNDTR_UPPER, -ndtr(-x), np.where(x > LOGNDTR_LOWER, mid, lower))


def dlog_ndtr(x):
    """Derivative TF's gradient of the selected ``where`` branch produces."""
    x = np.asarray(x, dtype=np.float64)
    xl = np.minimum(x, LOGNDTR_LOWER)
    x3 = xl * xl * xl
    dser = 2.0 / x3 - 12.0 / (x3 * xl * xl) + 90.0 / (x3 * x3 * xl)
    lower = (-xl - 1.0 / xl) + dser / _log_ndtr_series(xl)
    xm = np.maximum(x, LOGNDTR_LOWER)
    with np.errstate(divide='ignore', invalid='ignore'):
        mid = std_pdf(xm) / ndtr(xm)
    return np.where(x > LOGNDTR_UPPER, std_pdf(x), np.where(x > LOGNDTR_LOWER, mid, lower))


class MES:
    """mes.py:96-102: mean over the y* samples of gamma pdf(gamma) / (2 cdf(gamma)) - log cdf(gamma),
    gamma = (mean - y*) / sqrt(var) (no variance clamp in the reference)."""

    kind = 'mes'

    def __init__(self, gp, samples):
        self.gp, self.samples = gp, np.asarray(samples, dtype=np.float64).ravel()

    def gps(self):
        return [self.gp]

    def _terms(self, mean, var):
        gamma = (mean - self.samples[None, :]) / np.sqrt(var)
        pdf, cdf = std_pdf(gamma), ndtr(gamma)
        return gamma, pdf, cdf

    def value(self, X):
        mean, var = self.gp.predict(X)
        return self._value(mean[:, [0]], var[:, [0]])

    def _value(self, mean, var):
        gamma, pdf, cdf = self._terms(mean, var)
        return np.sum(gamma * pdf / (2.0 * cdf) - log_ndtr(gamma), axis=1, keepdims=True) / self.samples.size

    def value_and_grad(self, X):
        mean, var, dmean, dvar = self.gp.predict_with_grads(X)
        mean, var, dmean, dvar = mean[:, [0]], var[:, [0]], dmean[:, :, 0], dvar[:, :, 0]
        gamma, pdf, cdf = self._terms(mean, var)
        val = np.sum(gamma * pdf / (2.0 * cdf) - log_ndtr(gamma), axis=1, keepdims=True) / self.samples.size
        dA = (pdf - gamma * gamma * pdf) / (2.0 * cdf) - gamma * pdf * pdf / (2.0 * cdf * cdf)
        dh = dA - dlog_ndtr(gamma)                                   # d term / d gamma, N x S
        c_m = np.sum(dh, axis=1, keepdims=True) / np.sqrt(var)       # d gamma / d mean = 1 / s
        c_v = np.sum(dh * (-gamma / (2.0 * var)), axis=1, keepdims=True)
        return val, (c_m * dmean + c_v * dvar) / self.samples.size


# --------------------------------------------------------------------------------------------------
# acquisition tree
# --------------------------------------------------------------------------------------------------
class Leaf:
    """kind in {'ei','poi','pof','lcb'}; ``param`` is fmin / fmin / threshold / sigma."""

    def __init__(self, kind, gp, param):
        self.kind, self.gp, self.param = kind, gp, float(param)

    def gps(self):
        return [self.gp]

    def value(self, X):
        mean, var = self.gp.predict(X)
        return self._value(mean[:, [0]], var[:, [0]])

    def _value(self, mean, var):
        p = self.param
        if self.kind == 'lcb':                                  # lcb.py:55-57
            var = np.maximum(var, 0.0)
            return mean - p * np.sqrt(var)
        var = np.maximum(var, STABILITY)
        s = np.sqrt(var)
        if self.kind == 'ei':                                   # ei.py:72-79
            t1 = (p - mean) * normal_cdf(p, mean, s)
            t2 = var * normal_prob(p, mean, s)
            return t1 + t2
        if self.kind in ('poi', 'pof'):                         # poi.py:64-67, pof.py:82-85
            return normal_cdf(p, mean, s)
        raise ValueError(self.kind)

    def value_and_grad(self, X):
        mean, var, dmean, dvar = self.gp.predict_with_grads(X)
        mean, var, dmean, dvar = mean[:, [0]], var[:, [0]], dmean[:, :, 0], dvar[:, :, 0]
        p = self.param
        if self.kind == 'lcb':
            passes = (var >= 0.0)                               # tf.maximum: gradient to x iff x >= y
            v = np.maximum(var, 0.0)
            s = np.sqrt(v)
            with np.errstate(divide='ignore', invalid='ignore'):
                ds = np.where(passes, 1.0 / (2.0 * s), 0.0) * dvar
            return mean - p * s, dmean - p * ds
        passes = (var >= STABILITY)
        v = np.maximum(var, STABILITY)
        s = np.sqrt(v)
        ds = np.where(passes, 1.0 / (2.0 * s), 0.0) * dvar
        z = (p - mean) / s
        Phi, phi = ndtr(z), std_pdf(z)
        if self.kind == 'ei':
            val = (p - mean) * Phi + v * normal_prob(p, mean, s)
            return val, -Phi * dmean + phi * ds
        val = Phi
        return val, -(phi / s) * dmean - (z * phi / s) * ds


class Agg:
    """kind in {'sum','prod'}; optional constant ``scale`` (MCMC mean = sum / n, acquisition.py:446-448)."""

    def __init__(self, kind, operands, scale=1.0):
        self.kind, self.operands, self.scale = kind, list(operands), float(scale)

    def gps(self):
        return [g for o in self.operands for g in o.gps()]

    def value(self, X):
        cols = np.hstack([o.value(X) for o in self.operands])
        red = np.sum(cols, axis=1, keepdims=True) if self.kind == 'sum' else np.prod(cols, axis=1, keepdims=True)
        return red if self.scale == 1.0 else self.scale * red

    def value_and_grad(self, X):
        vg = [o.value_and_grad(X) for o in self.operands]
        cols = np.hstack([v for v, _ in vg])
        if self.kind == 'sum':
            val = np.sum(cols, axis=1, keepdims=True)
            grad = sum(g for _, g in vg)
        else:
            val = np.prod(cols, axis=1, keepdims=True)
            grad = 0.0
            for k, (_, g) in enumerate(vg):                     # zero-safe exclusive products
                others = np.prod(np.delete(cols, k, axis=1), axis=1, keepdims=True)
                grad = grad + others * g
        if self.scale != 1.0:
            val, grad = self.scale * val, self.scale * grad
        return val, grad


class HVPoI:
    """hvpoi.py:98-140 on top of a cell table produced by oracle.pareto.Pareto."""

    def __init__(self, gps, front, lb, ub, reference=None, chunk=64):
        self._gps = list(gps)
        self.front = np.asarray(front, dtype=np.float64)
        self.lb = np.asarray(lb, dtype=np.int64)
        self.ub = np.asarray(ub, dtype=np.int64)
        self.reference = estimate_reference(self.front) if reference is None else np.atleast_2d(reference)
        self.chunk = chunk

    def gps(self):
        return self._gps

    def pf_ext(self):
        R = self.front.shape[1]
        return np.vstack((-np.inf * np.ones((1, R)), self.front, self.reference))

    def _moments(self, X):
        preds = [g.predict(X) for g in self._gps]
        mean = np.hstack([p[0] for p in preds])
        var = np.hstack([p[1] for p in preds])
        return mean, np.maximum(var, STABILITY)

    def value(self, X):
        X = np.asarray(X, dtype=np.float64)
        out = np.empty((X.shape[0], 1))
        for s in range(0, X.shape[0], self.chunk):               # the literal graph holds N x C x R gathers
            mean, var = self._moments(X[s:s + self.chunk])
            out[s:s + self.chunk] = self._value_from_moments(mean, var)
        return out

    def _value_from_moments(self, mean, var):
        pf_ext = self.pf_ext()
        R = pf_ext.shape[1]
        col = np.arange(R)
        sd = np.sqrt(var)
        with np.errstate(invalid='ignore'):
            Phi = normal_cdf(pf_ext[None, :, :], mean[:, None, :], sd[:, None, :])       # N x (P+2) x R
        P1 = Phi[:, self.ub, col]                                                        # N x C x R
        P2 = Phi[:, self.lb, col]
        PoI = np.sum(np.prod(P1 - P2, axis=2), axis=1, keepdims=True)
        ub_points = pf_ext[self.ub, col]                                                 # C x R
        lb_points = pf_ext[self.lb, col]
        valid = np.all(ub_points[:, None, :] > mean[None, :, :], axis=2)                 # C x N
        splus_lb = np.maximum(lb_points[:, None, :], mean[None, :, :])                   # C x N x R
        splus = np.concatenate([valid[:, :, None].astype(np.float64),
                                ub_points[:, None, :] - splus_lb], axis=2)
        Hv = np.sum(np.prod(splus, axis=2), axis=0).reshape(-1, 1)
        return Hv * PoI

    def value_and_grad(self, X):
        """SURVEY.md Appendix A item 10 closed form (what tf.gradients yields)."""
        X = np.asarray(X, dtype=np.float64)
        N, D = X.shape
        val = np.empty((N, 1))
        grad = np.empty((N, D))
        pf_ext = self.pf_ext()
        R = pf_ext.shape[1]
        col = np.arange(R)
        ub_points, lb_points = pf_ext[self.ub, col], pf_ext[self.lb, col]
        for s in range(0, N, self.chunk):
            Xc = X[s:s + self.chunk]
            pr = [g.predict_with_grads(Xc) for g in self._gps]
            mean = np.hstack([p[0] for p in pr])
            var_raw = np.hstack([p[1] for p in pr])
            dmean = np.concatenate([p[2] for p in pr], axis=2)          # n x D x R
            dvar = np.concatenate([p[3] for p in pr], axis=2)
            passes = var_raw >= STABILITY
            var = np.maximum(var_raw, STABILITY)
            sd = np.sqrt(var)
            dsd = np.where(passes, 1.0 / (2.0 * sd), 0.0)[:, None, :] * dvar
            with np.errstate(invalid='ignore'):
                Z = (pf_ext[None, :, :] - mean[:, None, :]) / sd[:, None, :]      # n x (P+2) x R
                Phi = ndtr(Z)
                phi = np.where(np.isfinite(Z), std_pdf(Z), 0.0)
                zphi = np.where(np.isfinite(Z), Z * phi, 0.0)
            dP = Phi[:, self.ub, col] - Phi[:, self.lb, col]                       # n x C x R
            dphi = phi[:, self.ub, col] - phi[:, self.lb, col]
            dzphi = zphi[:, self.ub, col] - zphi[:, self.lb, col]
            PoI = np.sum(np.prod(dP, axis=2), axis=1)
            dPoI_dmu = np.zeros_like(mean)
            dPoI_dsd = np.zeros_like(mean)
            valid = np.all(ub_points[None, :, :] > mean[:, None, :], axis=2)       # n x C
            ext = ub_points[None, :, :] - np.maximum(lb_points[None, :, :], mean[:, None, :])
            Hv = np.sum(valid * np.prod(ext, axis=2), axis=1)
            dHv_dmu = np.zeros_like(mean)
            for j in range(R):
                others = np.prod(np.delete(dP, j, axis=2), axis=2)                 # n x C
                dPoI_dmu[:, j] = np.sum(others * (-dphi[:, :, j] / sd[:, [j]]), axis=1)
                dPoI_dsd[:, j] = np.sum(others * (-dzphi[:, :, j] / sd[:, [j]]), axis=1)
                oth_ext = np.prod(np.delete(ext, j, axis=2), axis=2)
                # tf.maximum(lb, mu): gradient goes to lb when lb >= mu, to mu otherwise
                moves = (lb_points[None, :, j] < mean[:, [j]])
                dHv_dmu[:, j] = -np.sum(valid * moves * oth_ext, axis=1)
            val[s:s + self.chunk, 0] = Hv * PoI
            g = np.zeros((Xc.shape[0], D))
            for j in range(R):
                g += (PoI * dHv_dmu[:, j] + Hv * dPoI_dmu[:, j])[:, None] * dmean[:, :, j]
                g += (Hv * dPoI_dsd[:, j])[:, None] * dsd[:, :, j]
            grad[s:s + self.chunk] = g
        return val, grad


def estimate_reference(front):
    """hvpoi.py:79-82."""
    f = np.max(front, axis=0, keepdims=True) - np.min(front, axis=0, keepdims=True)
    return np.max(front, axis=0, keepdims=True) + 2.0 * f / front.shape[0]


# --------------------------------------------------------------------------------------------------
# setup-side helpers and the candidate optimiser
# --------------------------------------------------------------------------------------------------
def fmin_of(gp, X_feasible):
    """ei.py:63-68 / poi.py:57-61: minimum posterior mean over the feasible training inputs (original units)."""
    mean, _ = gp.predict(X_feasible)
    return np.min(mean, axis=0)


def pof_feasible(gp, X, threshold=0.0, minimum_pof=0.5):
    """pof.py:63-79."""
    return Leaf('pof', gp, threshold).value(X).ravel() > minimum_pof


def mc_optimize(acq, points, chunk=4096):
    """optim.py:155-164 applied to inverse_acquisition (bo.py:244-245): argmin of -acq, first occurrence."""
    evaluations = np.vstack([-acq.value(points[s:s + chunk]) for s in range(0, points.shape[0], chunk)])
    idx_best = np.argmin(evaluations, axis=0)
    return points[idx_best, :], evaluations[idx_best, :], int(idx_best[0]), evaluations
