"""Forward kernel variants side by side: EI, D = 8, Matern52; M x variant grid, throughput (CUDA events around the contraction
launch) and oracle parity on a subset.   python tools/variant_sweep.py [variants, comma separated] [Ms, comma separated]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflowopt_b200 import _lib, synthetic           # noqa: E402
from oracle import acq as oacq, gp as ogp            # noqa: E402

variants = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "12,20,21").split(",")]
Ms = [int(m) for m in (sys.argv[2] if len(sys.argv) > 2 else "512,1024,2048,4096").split(",")]
D = 8
ctx = _lib.Context(0)
peak = ctx.measure_dmma_peak()
print("FP64 DMMA peak %.2f TF/s" % peak, flush=True)
for M in Ms:
    X, Y = synthetic.training_set(M, D)
    h = synthetic.hypers(D)
    g = ogp.ScaledGPR(X, Y, ogp.MATERN52, h["lengthscales"], h["variance"], h["noise_variance"])
    ctx.gp_set(0, g.Xs, g.Ys, _lib.KERN_MATERN52, **h)
    ctx.program_set([[_lib.OP_EI, 0, 0, 0]], [-1.5])
    N = int(min(4e6, 1e6 * (2048.0 / M) ** 2)) // 2
    Xc = synthetic.candidates(N, D)
    sub = np.random.RandomState(1).choice(N, 1500, replace=False)
    ref = oacq.Leaf('ei', g, -1.5).value(Xc[sub])
    for v in variants:
        try:
            ctx.set_variant(v)
        except _lib.GpfoError as e:
            print("variant %d: %s" % (v, e))
            continue
        ctx.eval(Xc[:20000], want_values=False)
        best = None
        for rep in range(2):
            vals, _, bv, bi = ctx.eval(Xc)
            t = ctx.timings()["hot_kernel_ms"]
            best = t if best is None else min(best, t)
        err = float(np.max(np.abs(vals[sub] - ref)) / np.max(np.abs(ref)))
        tf = N * M * (M + 1.0) / (best * 1e-3) / 1e12
        print("M=%5d variant %3d N=%8d: %8.2f ms  %.3e cand/s  %6.2f TF/s  %5.1f%% of DMMA peak  parity %.1e  argmax ok %s" % (
            M, v, N, best, N / (best * 1e-3), tf, 100 * tf / peak, err, bi == int(np.argmax(vals))), flush=True)
    ctx.set_variant(0)
