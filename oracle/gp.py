"""Oracle: GP regression prediction behind DataScaler, restated op-for-op in NumPy/SciPy float64.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  PARITY UNPINNED at 1e-9 (GPflow 0.5.0 / TF1 are
third-party dependencies absent from /root/reference; setup.py:33-34,70).

What is restated and where it comes from (paths relative to /root/reference):
  * LinearTransform forward / backward / backward_variance   gpflowopt/transforms.py:102-140, 155-157
  * Domain >> Domain  (A, b of the affine map)                 gpflowopt/domain.py:89-93
  * DataScaler X/Y setters and build_predict                   gpflowopt/scaling.py:165-190
  * GPR.build_predict                                          GPflow 0.5.0 gpr.py (called at
                                                               gpflowopt/scaling.py:189)
  * Stationary.square_dist / euclid_dist, RBF, Matern32/52     GPflow 0.5.0 kernels.py (models built at
                                                               testing/utility.py:47,55,61-62)
"""
import numpy as np
from scipy.linalg import solve_triangular, cholesky

RBF, MATERN52, MATERN32 = 0, 1, 2
KIND_NAMES = {RBF: "rbf", MATERN52: "matern52", MATERN32: "matern32"}


# --------------------------------------------------------------------------------------------------
# transforms.py / domain.py
# --------------------------------------------------------------------------------------------------
class Linear:
    """y = x A^T + b with A given as a vector (diagonal) or a matrix (transforms.py:82-100)."""

    def __init__(self, A, b):
        A = np.atleast_1d(np.asarray(A, dtype=np.float64))
        self.A = np.diag(A) if A.ndim == 1 else A
        self.b = np.atleast_1d(np.asarray(b, dtype=np.float64))

    def forward(self, X):                       # transforms.py:102-103
        return np.matmul(X, self.A.T) + self.b

    def backward(self, Y):                      # transforms.py:112-118 (cholesky of A^T, cholesky_solve)
        L = np.linalg.cholesky(self.A.T)
        Z = solve_triangular(L, (Y - self.b).T, lower=True)
        return solve_triangular(L.T, Z, lower=False).T

    def backward_variance_literal(self, Yvar):  # transforms.py:120-140, N x R input, N x N x R temporary
        N, R = Yvar.shape
        full = np.zeros((N, N, R))
        idx = np.arange(N)
        full[idx, idx, :] = Yvar                # matrix_diag(transpose) then transpose(perm=[1,2,0])
        L = np.linalg.cholesky(np.square(self.A.T))
        flat = full.reshape(N * N, R)
        Z = solve_triangular(L, flat.T, lower=True)
        scaled = solve_triangular(L.T, Z, lower=False).T.reshape(N, N, R)
        return scaled.sum(axis=1)

    def backward_variance(self, Yvar):
        """Algebraically identical O(N) form for diagonal A (what the CUDA epilogue implements)."""
        a = np.diag(self.A)
        assert np.array_equal(self.A, np.diag(a)), "O(N) form needs a diagonal transform"
        return Yvar / np.square(a)

    def invert(self):                           # transforms.py:155-157
        A_inv = np.linalg.inv(self.A.T)
        return Linear(A_inv, -np.dot(self.b, A_inv))


def domain_transform(lower1, upper1, lower2, upper2):
    """Affine map of box [lower1, upper1] onto [lower2, upper2] (domain.py:89-93)."""
    lower1, upper1, lower2, upper2 = (np.asarray(v, dtype=np.float64) for v in (lower1, upper1, lower2, upper2))
    A = (upper2 - lower2) / (upper1 - lower1)
    b = -upper1 * A + upper2
    return Linear(A, b)


# --------------------------------------------------------------------------------------------------
# GPflow 0.5.0 kernels (recalled semantics)
# --------------------------------------------------------------------------------------------------
def square_dist(X, X2, lengthscales):
    X = X / lengthscales
    Xs = np.sum(np.square(X), axis=1)
    if X2 is None:
        return -2.0 * np.matmul(X, X.T) + Xs.reshape(-1, 1) + Xs.reshape(1, -1)
    X2 = X2 / lengthscales
    X2s = np.sum(np.square(X2), axis=1)
    return -2.0 * np.matmul(X, X2.T) + Xs.reshape(-1, 1) + X2s.reshape(1, -1)


def kern_K(kind, variance, lengthscales, X, X2=None):
    r2 = square_dist(X, X2, lengthscales)
    if kind == RBF:
        return variance * np.exp(-r2 / 2.0)
    r = np.sqrt(r2 + 1e-12)
    if kind == MATERN52:
        return variance * (1.0 + np.sqrt(5.0) * r + 5.0 / 3.0 * np.square(r)) * np.exp(-np.sqrt(5.0) * r)
    if kind == MATERN32:
        return variance * (1.0 + np.sqrt(3.0) * r) * np.exp(-np.sqrt(3.0) * r)
    raise ValueError(kind)


def kern_dK_dr2(kind, variance, r2):
    """d k / d r^2 (needed by the closed-form gradient)."""
    if kind == RBF:
        return -0.5 * variance * np.exp(-r2 / 2.0)
    r = np.sqrt(r2 + 1e-12)
    if kind == MATERN52:
        return -(5.0 / 6.0) * variance * (1.0 + np.sqrt(5.0) * r) * np.exp(-np.sqrt(5.0) * r)
    if kind == MATERN32:
        return -1.5 * variance * np.exp(-np.sqrt(3.0) * r)
    raise ValueError(kind)


# --------------------------------------------------------------------------------------------------
# GPR behind a DataScaler
# --------------------------------------------------------------------------------------------------
class ScaledGPR:
    """One GPR model wrapped by a DataScaler (gpflowopt/scaling.py).

    ``X``, ``Y`` are given in ORIGINAL units; the scaler stores transformed copies in the wrapped model
    exactly as the X/Y setters do (scaling.py:165-181).  Hyper-parameters act on the scaled data.
    """

    def __init__(self, X, Y, kind, lengthscales, variance, noise_variance,
                 domain_lower=None, domain_upper=None, normalize_Y=False):
        X = np.asarray(X, dtype=np.float64)
        Y = np.asarray(Y, dtype=np.float64)
        D, R = X.shape[1], Y.shape[1]
        if domain_lower is None:                 # identity: UnitCube >> UnitCube (scaling.py:70)
            self.input_transform = domain_transform(np.zeros(D), np.ones(D), np.zeros(D), np.ones(D))
        else:                                    # domain >> UnitCube (acquisition.py:141)
            self.input_transform = domain_transform(domain_lower, domain_upper, np.zeros(D), np.ones(D))
        if normalize_Y:                          # scaling.py:179-180: ~LinearTransform(std, mean), ddof=0
            self.output_transform = Linear(Y.std(axis=0), Y.mean(axis=0)).invert()
        else:
            self.output_transform = Linear(np.ones(R), np.zeros(R))
        self.Xs = self.input_transform.forward(X)
        self.Ys = self.output_transform.forward(Y)
        self.kind = kind
        self.lengthscales = np.asarray(lengthscales, dtype=np.float64)   # scalar or length D (ARD)
        self.variance = float(variance)
        self.noise_variance = float(noise_variance)
        self._factor = None

    # -- GPflow gpr.py: K = kern.K(X) + eye * likelihood.variance; L = cholesky(K); V = L^-1 Y
    def factor(self):
        if self._factor is None:
            K = kern_K(self.kind, self.variance, self.lengthscales, self.Xs)
            K = K + np.eye(self.Xs.shape[0]) * self.noise_variance
            L = cholesky(K, lower=True)
            V = solve_triangular(L, self.Ys, lower=True)
            self._factor = (L, V)
        return self._factor

    def predict_scaled(self, Xnew_scaled):
        """GPR.build_predict(Xnew, full_cov=False) on already input-transformed points."""
        L, V = self.factor()
        Kx = kern_K(self.kind, self.variance, self.lengthscales, self.Xs, Xnew_scaled)
        A = solve_triangular(L, Kx, lower=True)
        fmean = np.matmul(A.T, V)
        fvar = self.variance - np.sum(np.square(A), axis=0)
        fvar = np.tile(fvar.reshape(-1, 1), (1, self.Ys.shape[1]))
        return fmean, fvar

    def predict(self, Xcand, literal_variance=False):
        """DataScaler.build_predict (scaling.py:183-190): forward inputs, predict, un-scale outputs."""
        f, var = self.predict_scaled(self.input_transform.forward(np.asarray(Xcand, dtype=np.float64)))
        mean = self.output_transform.backward(f)
        if literal_variance:
            return mean, self.output_transform.backward_variance_literal(var)
        return mean, self.output_transform.backward_variance(var)

    # -- closed-form gradients of (mean, var) w.r.t. the ORIGINAL candidate coordinates (row J)
    def predict_with_grads(self, Xcand):
        Xcand = np.asarray(Xcand, dtype=np.float64)
        L, V = self.factor()
        a_in = np.diag(self.input_transform.A)
        ell = np.broadcast_to(self.lengthscales, (Xcand.shape[1],))
        Xn = self.input_transform.forward(Xcand)
        r2 = square_dist(self.Xs, Xn, self.lengthscales)                       # M x N
        Kx = kern_K(self.kind, self.variance, self.lengthscales, self.Xs, Xn)
        dk = kern_dK_dr2(self.kind, self.variance, r2)                          # M x N
        A = solve_triangular(L, Kx, lower=True)
        W = solve_triangular(L.T, A, lower=False)                               # K^-1 kx, M x N
        alpha = solve_triangular(L.T, V, lower=False)                           # M x R
        fmean = np.matmul(A.T, V)
        fvar = self.variance - np.sum(np.square(A), axis=0)
        # d r2[i,n] / d x[n,d] = 2 (xn[n,d] - Xs[i,d]) / ell_d^2 * a_in_d
        diff = (Xn[None, :, :] - self.Xs[:, None, :]) / (ell * ell) * (2.0 * a_in)   # M x N x D
        a_out = np.diag(self.output_transform.A)
        R = self.Ys.shape[1]
        dmean = np.stack([np.einsum('i,in,ind->nd', alpha[:, r], dk, diff) / a_out[r] for r in range(R)], axis=2)
        dvar_s = -2.0 * np.einsum('in,in,ind->nd', W, dk, diff)                  # N x D (scaled units)
        dvar = np.stack([dvar_s / (a_out[r] ** 2) for r in range(R)], axis=2)    # N x D x R
        mean = self.output_transform.backward(fmean)
        var = self.output_transform.backward_variance(np.tile(fvar.reshape(-1, 1), (1, R)))
        return mean, var, dmean, dvar
