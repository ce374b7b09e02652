"""Host-side mirror of the reference's data transforms (gpflowopt/transforms.py).

Only what the acquisition path needs: an affine map y = x A^T + b, its inverse and the variance scaling.  The
transforms are plain NumPy on the host; on the device they are folded into the contraction kernel's prologue
(input map) and epilogue (output un-scaling), which is why only DIAGONAL maps are accepted by the engine.
"""
import numpy as np


class DataTransform(object):
    """Maps data in Domain U to Domain V (gpflowopt/transforms.py:24-71)."""

    def forward(self, X):
        raise NotImplementedError

    def backward(self, Y):
        return (~self).forward(Y)

    def assign(self, other):
        raise NotImplementedError

    def __invert__(self):
        raise NotImplementedError


class LinearTransform(DataTransform):
    """Y = (A X^T)^T + b  (gpflowopt/transforms.py:74-157).  ``A`` may be a vector (diagonal) or a matrix."""

    def __init__(self, A, b):
        assert A is not None and b is not None
        b = np.atleast_1d(np.asarray(b, dtype=np.float64))
        A = np.atleast_1d(np.asarray(A, dtype=np.float64))
        if A.ndim == 1:
            A = np.diag(A)
        assert b.ndim == 1 and A.ndim == 2
        self.A = A
        self.b = b
        self.version = 0

    def forward(self, X):
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        diag = np.diagonal(self.A)
        if X.shape[0] > 4096 and np.count_nonzero(self.A) == np.count_nonzero(diag) and np.all(np.isfinite(X)):
            return X * diag + self.b             # diagonal map: the products the matmul forms, without the D x D zeros
        return np.matmul(X, self.A.T) + self.b

    def backward(self, Y):
        Y = np.atleast_2d(np.asarray(Y, dtype=np.float64))
        return np.linalg.solve(self.A, (Y - self.b).T).T

    def backward_variance(self, Yvar):
        """N x P diagonal variances (gpflowopt/transforms.py:120-140 for rank-2 input)."""
        Yvar = np.atleast_2d(np.asarray(Yvar, dtype=np.float64))
        return np.linalg.solve(np.square(self.A), Yvar.T).T

    def assign(self, other):
        assert isinstance(other, LinearTransform)
        self.A = np.array(other.A, dtype=np.float64)
        self.b = np.array(other.b, dtype=np.float64)
        self.version += 1

    def __invert__(self):
        A_inv = np.linalg.inv(self.A.T)
        return LinearTransform(A_inv, -np.dot(self.b, A_inv))

    def is_diagonal(self):
        return bool(np.array_equal(self.A, np.diag(np.diag(self.A))))

    def diagonal(self):
        """(scale, shift) vectors of a diagonal map; raises for a general matrix (outside the device path)."""
        if not self.is_diagonal():
            raise ValueError("only diagonal LinearTransforms are supported by the device path")
        return np.diag(self.A).copy(), self.b.copy()

    def __str__(self):
        return "XA + b"
