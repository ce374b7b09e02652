"""Repeats small-batch evaluations (row-split and un-split paths, values / gradients / predict) many times and compares every
result bit for bit with the first repetition: any difference means a race or an uninitialised read."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflowopt_b200 import _lib, synthetic           # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ctx = _lib.Context(0)
bad = 0
for M, D in ((2048, 8), (64, 4), (300, 3), (1100, 8)):
    X, Y = synthetic.training_set(M, D)
    ctx.gp_set(0, X, Y, _lib.KERN_MATERN52, **synthetic.hypers(D))
    ctx.program_set([[_lib.OP_EI, 0, 0, 0]], [-1.0])
    if M == 2048:                                      # leave large stale buffers behind, like a full test session
        ctx.eval(synthetic.candidates(300000, D), want_grads=True)
        continue
    first = {}
    for rep in range(reps):
        for N in (1, 7, 8, 9, 100, 592, 593, 1500):
            Xc = synthetic.candidates(N, D, seed=N)
            for variant in (0, 6):
                ctx.set_variant(variant)
                v, g, bv, bi = ctx.eval(Xc, want_grads=True)
                v2, _, bv2, bi2 = ctx.eval(Xc)
                m, s = ctx.predict(0, Xc, 1)
                key = (N, variant)
                cur = (v.copy(), g.copy(), bv, bi, v2.copy(), bv2, bi2, m.copy(), s.copy())
                if key not in first:
                    first[key] = cur
                else:
                    names = ("values(grad call)", "grads", "best value", "best index", "values", "best value 2", "best index 2", "mean", "var")
                    for nm, a, b in zip(names, first[key], cur):
                        same = np.array_equal(a, b) if isinstance(a, np.ndarray) else a == b
                        if not same:
                            bad += 1
                            if bad < 20:
                                d = np.max(np.abs(np.asarray(a) - np.asarray(b))) if isinstance(a, np.ndarray) else (a, b)
                                print("MISMATCH M=%d N=%d variant=%d rep=%d %s: %r" % (M, N, variant, rep, nm, d), flush=True)
            ctx.set_variant(0)
    print("M=%d: %d repetitions done, mismatches so far %d" % (M, reps, bad), flush=True)
print("TOTAL MISMATCHES", bad)
