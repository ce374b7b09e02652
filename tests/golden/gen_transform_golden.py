"""Generate tests/golden/transform_golden.npz by running the REFERENCE's own LinearTransform graph builders and domain maps
(transforms.py:74-157, domain.py:89-93) on fixed data.

Run in the build container only (needs /root/reference):  python tests/golden/gen_transform_golden.py

Executed reference code: ``Domain.__rshift__``, ``LinearTransform.__init__ / __invert__ / build_forward / build_backward /
build_backward_variance`` (the N x N x R path the reference takes for diagonal variances).  TensorFlow primitives are NumPy /
SciPy calls of the same meaning; ``tf.cholesky_solve`` is two triangular solves, as in TensorFlow.
"""
import importlib
import importlib.util
import os
import sys

import numpy as np
from scipy.linalg import solve_triangular

HERE = os.path.dirname(os.path.abspath(__file__))


def numpy_tf(tf):
    def matrix_diag(x):                                   # (..., N) -> (..., N, N)
        x = np.asarray(x)
        out = np.zeros(x.shape + (x.shape[-1],))
        idx = np.arange(x.shape[-1])
        out[..., idx, idx] = x
        return out

    def cholesky_solve(L, rhs):
        return solve_triangular(L, solve_triangular(L, rhs, lower=True), lower=True, trans='T')

    tf.matmul, tf.square, tf.cholesky = np.matmul, np.square, np.linalg.cholesky
    tf.transpose = lambda x, perm=None: np.transpose(x, perm)
    tf.rank = lambda x: np.ndim(x)
    tf.equal = lambda a, b: a == b
    tf.cond = lambda pred, fn1, fn2: fn1() if pred else fn2()
    tf.matrix_diag, tf.cholesky_solve = matrix_diag, cholesky_solve
    tf.shape = lambda x: np.array(np.shape(x))
    tf.reshape = lambda x, shape: np.reshape(x, [int(s) for s in shape])
    tf.reduce_sum = lambda x, axis=None, keep_dims=False: np.sum(x, axis=axis, keepdims=keep_dims)


def main():
    spec = importlib.util.spec_from_file_location("gen_mes_golden", os.path.join(HERE, "gen_mes_golden.py"))
    stubs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(stubs)
    stubs._install_stubs()
    numpy_tf(sys.modules["tensorflow"])
    domain = importlib.import_module("gpflowopt.domain")
    transforms = importlib.import_module("gpflowopt.transforms")
    rs = np.random.RandomState(20240615)
    out = {}
    for k, (d, r, n) in enumerate(((2, 1, 20), (5, 2, 30), (8, 1, 25))):
        lo, up = rs.randn(d) * 4 - 2, None
        up = lo + rs.rand(d) * 12 + 0.5
        box = np.sum([domain.ContinuousParameter("x%d" % i, l, u) for i, (l, u) in enumerate(zip(lo, up))])
        cube = np.sum([domain.ContinuousParameter("u%d" % i, 0, 1) for i in range(d)])
        t_in = box >> cube
        X = lo + (up - lo) * rs.rand(n, d)
        Y = rs.randn(n, r) * np.array([3.0, 0.2][:r]) + np.array([2.0, -7.0][:r])
        t_out = ~transforms.LinearTransform(Y.std(axis=0), Y.mean(axis=0))          # scaling.py:179-180
        f_scaled, v_scaled = rs.randn(n, r), np.abs(rs.randn(n, r)) * 0.3
        out.update({"c%d_lower" % k: lo, "c%d_upper" % k: up, "c%d_X" % k: X, "c%d_Y" % k: Y,
                    "c%d_f_scaled" % k: f_scaled, "c%d_v_scaled" % k: v_scaled,
                    "c%d_in_A" % k: np.array(t_in.A), "c%d_in_b" % k: np.array(t_in.b),
                    "c%d_out_A" % k: np.array(t_out.A), "c%d_out_b" % k: np.array(t_out.b),
                    "c%d_X_scaled" % k: np.array(t_in.build_forward(X)), "c%d_Y_scaled" % k: np.array(t_out.build_forward(Y)),
                    "c%d_X_back" % k: np.array(t_in.build_backward(np.array(t_in.build_forward(X)))),
                    "c%d_mean" % k: np.array(t_out.build_backward(f_scaled)),
                    "c%d_var" % k: np.array(t_out.build_backward_variance(v_scaled))})
        print("case", k, "var[:2]", out["c%d_var" % k][:2].ravel(), "mean[:2]", out["c%d_mean" % k][:2].ravel())
    out["n_cases"] = np.array(3)
    np.savez(os.path.join(HERE, "transform_golden.npz"), **out)


if __name__ == "__main__":
    main()
