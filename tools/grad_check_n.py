import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from gpflowopt_b200 import _lib, synthetic
ctx = _lib.Context(0)
peak = ctx.measure_dmma_peak()
for M in (4096,):
    X, Y = synthetic.training_set(M, 8)
    ctx.gp_set(0, X, Y, _lib.KERN_MATERN52, **synthetic.hypers(8))
    ctx.program_set([[_lib.OP_EI, 0, 0, 0]], [-1.5])
    for N in (40000, 50000, 62500, 125000):
        Xc = synthetic.candidates(N, 8)
        ctx.eval(Xc, want_grads=True)
        best = 1e30
        for _ in range(3):
            ctx.eval(Xc, want_grads=True); best = min(best, ctx.timings()["eval_ms"])
        tf = N * 2.0 * M * (M + 1.0) / (best * 1e-3) / 1e12
        print("M=%d N=%d: %.1f ms %.1f%% launches %d" % (M, N, best, 100 * tf / peak, ctx.timings()["launches"]), flush=True)
