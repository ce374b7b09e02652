"""Mid-size batches (a few hundred to a few ten thousand candidates): automatic kernel choice against pinned variants.
   python tools/midsize.py [M]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflowopt_b200 import _lib, synthetic
M = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
D = 8
ctx = _lib.Context(0)
X, Y = synthetic.training_set(M, D)
ctx.gp_set(0, X, Y, _lib.KERN_MATERN52, **synthetic.hypers(D))
ctx.program_set([[_lib.OP_EI, 0, 0, 0]], [-1.5])
for N in [int(a) for a in sys.argv[2].split(',')] if len(sys.argv) > 2 else (1000, 2368, 3000, 5000, 10000, 20000, 30000, 37888):
    Xc = synthetic.candidates(N, D)
    row = []
    ref = None
    for v in (0, 6, 5, 11, 12, 19, 20, 21):
        ctx.set_variant(v)
        ctx.eval(Xc)
        best = 1e30
        for _ in range(5):
            vals, _, bv, bi = ctx.eval(Xc)
            best = min(best, ctx.timings()["eval_ms"])
        if ref is None:
            ref = vals
            chosen = ctx.timings()["variant"]
        else:
            assert np.max(np.abs(vals - ref)) <= 1e-11 * np.max(np.abs(ref)), (N, v)
        row.append("%s %.3f" % ("auto(%d)" % chosen if v == 0 else "v%d" % v, best))
    ctx.set_variant(0)
    print("M=%d N=%6d ms: " % (M, N) + "  ".join(row), flush=True)
