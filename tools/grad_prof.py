"""One gradient evaluation of the headline model for ncu:  python tools/grad_prof.py [N] [M]
(contract2_kernel launches in order: forward with stored A, backward over the reversed transposed stream)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflowopt_b200 import _lib, synthetic           # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
D = 8
X, Y = synthetic.training_set(M, D)
ctx = _lib.Context(0)
ctx.gp_set(0, X, Y, _lib.KERN_MATERN52, **synthetic.hypers(D))
ctx.program_set([[_lib.OP_EI, 0, 0, 0]], [-1.5])
Xc = synthetic.candidates(N, D)
ctx.eval(Xc, want_grads=True)
t = ctx.timings()
print("M=%d N=%d gradients: eval %.3f ms -> %.3e cand/s, %.2f TF/s algorithmic, %d launches" % (
    M, N, t["eval_ms"], N / (t["eval_ms"] * 1e-3), N * 2.0 * M * (M + 1.0) / (t["eval_ms"] * 1e-3) / 1e12, t["launches"]))
