"""One N = 1 gradient call at M = 2048 for an ncu launch list (per-kernel durations of the small-call path)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflowopt_b200 import _lib, synthetic
M = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
D = 8
X, Y = synthetic.training_set(M, D)
ctx = _lib.Context(0)
ctx.gp_set(0, X, Y, _lib.KERN_MATERN52, **synthetic.hypers(D))
ctx.program_set([[_lib.OP_EI, 0, 0, 0]], [-1.5])
Xc = synthetic.candidates(1, D)
for _ in range(3):
    ctx.eval(Xc, want_grads=True)
print(ctx.timings())
