"""Gradient-pass throughput on one B200: stored-A path (default for large batches) vs the dense K^-1 pass (GPFO_GRAD_DENSE=1).
Algorithmic work per candidate: 2 M (M + 1) flop (F_grad).  Writes to stdout; run under gpurun."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflowopt_b200 import _lib, synthetic                    # noqa: E402
from oracle import acq as oacq, gp as ogp                     # noqa: E402


def main():
    ctx = _lib.Context(0)
    peak = ctx.measure_dmma_peak()
    print("FP64 DMMA peak %.2f TF/s" % peak, flush=True)
    for M in (512, 1024, 2048, 4096, 8192):
        X, Y = synthetic.training_set(M, 8, seed=1234)
        h = synthetic.hypers(8)
        g = ogp.ScaledGPR(X, Y[:, [0]], ogp.MATERN52, h["lengthscales"], h["variance"], h["noise_variance"])
        ctx.gp_set(0, g.Xs, g.Ys, g.kind, g.lengthscales, g.variance, g.noise_variance, np.diag(g.input_transform.A),
                   g.input_transform.b, np.diag(g.output_transform.A), g.output_transform.b, 0)
        ctx.gp_count_set(1)
        ctx.program_set([[_lib.OP_EI, 0, 0, 0]], [-1.5])
        N = int(max(40000, min(1e6, 2.5e5 * (2048.0 / M) ** 2)))
        Xc = synthetic.candidates(N, 8)
        _, rg = oacq.Leaf('ei', g, -1.5).value_and_grad(Xc[:200])
        for mode in ("stored-A", "dense"):
            if mode == "dense":
                os.environ["GPFO_GRAD_DENSE"] = "1"
            else:
                os.environ.pop("GPFO_GRAD_DENSE", None)
            ctx.eval(Xc, want_grads=True)
            best = 1e30
            for _ in range(2):
                vals, grads, _, _ = ctx.eval(Xc, want_grads=True)
                best = min(best, ctx.timings()["eval_ms"])
            err = float(np.max(np.abs(grads[:200] - rg)) / np.max(np.abs(rg)))
            tf = N * 2.0 * M * (M + 1.0) / (best * 1e-3) / 1e12
            print("M=%5d N=%8d %-8s: %8.1f ms -> %.3e cand/s, %.2f TF/s algorithmic (%.1f%% of peak), %d launches, grad parity %.1e"
                  % (M, N, mode, best, N / (best * 1e-3), tf, 100 * tf / peak, ctx.timings()["launches"], err), flush=True)
    os.environ.pop("GPFO_GRAD_DENSE", None)


if __name__ == "__main__":
    main()
