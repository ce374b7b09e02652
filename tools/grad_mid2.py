"""Mid-size gradient batches: dense K^-1 pass (3 M^2 flop, 32-candidate tiles, wave-filling row split) against the stored-A pass
(2 M^2 flop; forward on 16-candidate tiles, backward on 64-candidate tiles) -- where does the stored-A pass start to win?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflowopt_b200 import _lib, synthetic           # noqa: E402
from oracle import acq as oacq, gp as ogp            # noqa: E402
ctx = _lib.Context(0)
for M in (1024, 2048, 4096):
    X, Y = synthetic.training_set(M, 8)
    h = synthetic.hypers(8)
    g = ogp.ScaledGPR(X, Y, ogp.MATERN52, h["lengthscales"], h["variance"], h["noise_variance"])
    ctx.gp_set(0, g.Xs, g.Ys, _lib.KERN_MATERN52, **h)
    ctx.program_set([[_lib.OP_EI, 0, 0, 0]], [-1.5])
    for N in (2400, 3000, 4736, 6000, 9472, 12000, 20000, 30000):
        Xc = synthetic.candidates(N, 8)
        _, rg = oacq.Leaf('ei', g, -1.5).value_and_grad(Xc[:64])
        out = []
        for min_n in ("1000000000", "1"):
            os.environ["GPFO_ASTORE_MIN_N"] = min_n
            ctx.eval(Xc, want_grads=True)
            best = 1e30
            for _ in range(3):
                _, grads, _, _ = ctx.eval(Xc, want_grads=True)
                best = min(best, ctx.timings()["eval_ms"])
            out.append((best, float(np.max(np.abs(grads[:64] - rg)) / np.max(np.abs(rg)))))
        print("M=%d N=%6d: dense pass %.3f ms (parity %.1e) | stored-A pass %.3f ms (parity %.1e)  x%.2f" % (
            M, N, out[0][0], out[0][1], out[1][0], out[1][1], out[0][0] / out[1][0]), flush=True)
os.environ.pop("GPFO_ASTORE_MIN_N", None)
