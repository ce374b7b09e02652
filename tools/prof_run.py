"""One short evaluation of the headline shape for ncu:  python tools/prof_run.py [variant] [N] [M]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflowopt_b200 import _lib, synthetic           # noqa: E402

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
N = int(sys.argv[2]) if len(sys.argv) > 2 else 9472
M = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
D = 8
X, Y = synthetic.training_set(M, D)
ctx = _lib.Context(0)
ctx.gp_set(0, X, Y, _lib.KERN_MATERN52, **synthetic.hypers(D))
ctx.program_set([[_lib.OP_EI, 0, 0, 0]], [-1.5])
ctx.set_variant(variant)
Xc = synthetic.candidates(N, D)
for _ in range(reps):
    ctx.eval(Xc, want_values=False)
    t = ctx.timings()
    print("variant %d M=%d N=%d: hot %.3f ms -> %.3e cand/s %.2f TF/s" % (
        variant, M, N, t["hot_kernel_ms"], N / (t["hot_kernel_ms"] * 1e-3), N * M * (M + 1.0) / (t["hot_kernel_ms"] * 1e-3) / 1e12))
