"""First-contact script for a GPU box: factor / predict / EI parity vs the oracle on small sizes, then timing of
the contraction kernel variants and the DMMA peak.  Writes gpurun_out/smoke.log."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflowopt_b200 import _lib, synthetic           # noqa: E402
from oracle import acq as oacq, gp as ogp            # noqa: E402


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


def main():
    ctx = _lib.Context(0)
    print("dmma peak TF/s:", ctx.measure_dmma_peak(), flush=True)
    for (M, D, kind, noise) in [(20, 2, "rbf", 1e-2), (100, 3, "matern52", 1e-2), (515, 8, "matern52", 1e-3),
                                (1100, 8, "rbf", 1e-2), (2048, 8, "matern52", 1e-2)]:
        X, Y = synthetic.training_set(M, D)
        h = synthetic.hypers(D, noise)
        g = ogp.ScaledGPR(X, Y, synthetic.KIND[kind], h["lengthscales"], h["variance"], h["noise_variance"])
        t0 = time.time()
        ctx.gp_set(0, g.Xs, g.Ys, synthetic.KIND[kind], **h)
        t_set = time.time() - t0
        L, Linv, V = ctx.gp_get_factor(0, M, 1)
        Lo, Vo = g.factor()
        print("M=%d D=%d %s: gp_set %.3fs factor_ms=%.2f  L rel %.2e  V rel %.2e  Linv*L-I %.2e" % (
            M, D, kind, t_set, ctx.timings()["factor_ms"], rel(L, Lo), rel(V, Vo),
            float(np.max(np.abs(Linv @ Lo - np.eye(M))))), flush=True)
        for N in (1, 7, 16, 17, 1000, 5000):
            Xc = synthetic.candidates(N, D)
            mean, var = ctx.predict(0, Xc, 1)
            mo, vo = g.predict(Xc)
            fmin = float(np.min(g.predict(X)[0]))
            ctx.program_set([[_lib.OP_EI, 0, 0, 0]], [fmin])
            for variant in (12, 4):
                ctx.set_variant(variant)
                vals, _, bv, bi = ctx.eval(Xc)
                ref = oacq.Leaf('ei', g, fmin).value(Xc)
                print("   N=%d v%d: mean rel %.2e var rel %.2e  EI scale-rel %.2e  argmax %s (%d vs %d)" % (
                    N, variant, rel(mean, mo), float(np.max(np.abs(var - vo)) / h["variance"]), rel(vals, ref),
                    bi == int(np.argmax(ref)), bi, int(np.argmax(ref))), flush=True)
            ctx.set_variant(0)
    # timing at the headline shape
    M, D = 2048, 8
    for N in (1000000,):
        Xc = synthetic.candidates(N, D)
        for variant in (12, 4):
            ctx.set_variant(variant)
            for rep in range(2):
                ctx.eval(Xc, want_values=False)
                t = ctx.timings()
                print("M=%d N=%d variant %d: eval_ms %.2f hot %.2f h2d %.2f -> %.3e cand/s, %.2f TF/s" % (
                    M, N, variant, t["eval_ms"], t["hot_kernel_ms"], t["h2d_ms"], N / (t["eval_ms"] * 1e-3),
                    N * M * (M + 1.0) / (t["hot_kernel_ms"] * 1e-3) / 1e12), flush=True)


if __name__ == "__main__":
    main()
