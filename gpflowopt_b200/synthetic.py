"""Seeded synthetic workloads of SURVEY.md section 8(d), shared by bench.py, smoke() and the tests so that the
CUDA path and the CPU oracle are fed identical bytes."""
import numpy as np

from . import _lib


def training_set(M, D, seed=1234, constraint=False):
    """X ~ U[0,1]^D, y = sin(3 sum x) + 0.1 eps standardised (and optionally a linear constraint output)."""
    rs = np.random.RandomState(seed)
    X = rs.rand(M, D)
    y = np.sin(3.0 * X.sum(axis=1)) + 0.1 * rs.randn(M)
    y = (y - y.mean()) / (y.std() if y.std() > 0 else 1.0) + (1.0 if M == 1 else 0.0)
    Y = y.reshape(-1, 1)
    if constraint:
        yc = X[:, 0] - 0.5 + 0.05 * rs.randn(M)
        Y = np.hstack((Y, yc.reshape(-1, 1)))
    return X, Y


def hypers(D, noise=1e-2):
    """Fixed hyper-parameters (no MLE): ARD lengthscale 0.7 sqrt(D/8), kernel variance 1.3."""
    return dict(lengthscales=np.full(D, 0.7 * np.sqrt(D / 8.0)), variance=1.3, noise_variance=noise)


def candidates(N, D, seed=4321, lower=None, upper=None):
    X = np.random.RandomState(seed).rand(N, D)
    if lower is not None:
        X = np.asarray(lower) + X * (np.asarray(upper) - np.asarray(lower))
    return X


def dtlz2_objectives(X, R=3, g_scale=1.0):
    """DTLZ2-style objectives on the first R-1 coordinates (config #4).  ``g_scale`` weighs the distance term of the remaining
    coordinates: 0.4 gives a Pareto front of ~200 of 1024 uniform points at D = 6 (SURVEY 8d asks for P = 200 +- 20)."""
    g = g_scale * np.sum((X[:, R - 1:] - 0.5) ** 2, axis=1)
    theta = X[:, :R - 1] * (np.pi / 2.0)
    F = np.empty((X.shape[0], R))
    for j in range(R):
        f = 1.0 + g
        for t in range(R - 1 - j):
            f = f * np.cos(theta[:, t])
        if j > 0:
            f = f * np.sin(theta[:, R - 1 - j])
        F[:, j] = f
    return F


KIND = {"rbf": _lib.KERN_RBF, "matern52": _lib.KERN_MATERN52, "matern32": _lib.KERN_MATERN32}
