"""Mid-size gradient batches (dense K^-1 pass): wave-filling row split vs one CTA per tile."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflowopt_b200 import _lib, synthetic           # noqa: E402
ctx = _lib.Context(0)
for M in (2048, 4096):
    X, Y = synthetic.training_set(M, 8)
    ctx.gp_set(0, X, Y, _lib.KERN_MATERN52, **synthetic.hypers(8))
    ctx.program_set([[_lib.OP_EI, 0, 0, 0]], [-1.5])
    for N in (1000, 2400, 3000, 5000, 8000, 12000, 20000, 30000):
        Xc = synthetic.candidates(N, 8)
        out = []
        for off in ("1", "0"):
            os.environ["GPFO_GRAD_NO_WAVE_SPLIT"] = off
            ctx.eval(Xc, want_grads=True)
            best = min(ctx.eval(Xc, want_grads=True) and ctx.timings()["eval_ms"] for _ in range(3))
            out.append((best, ctx.timings()["launches"]))
        print("M=%d N=%6d: one CTA per tile %.3f ms (%d launches) -> wave-filling split %.3f ms (%d launches)  x%.2f" % (
            M, N, out[0][0], out[0][1], out[1][0], out[1][1], out[0][0] / out[1][0]), flush=True)
os.environ.pop("GPFO_GRAD_NO_WAVE_SPLIT", None)
