"""Per-warp phase cycle counters of the profiling build (variant 127) of the default kernel."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflowopt_b200 import _lib, synthetic
M = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
D, N = 8, int(1000000 * min(4.0, (2048.0 / M) ** 2))
X, Y = synthetic.training_set(M, D)
ctx = _lib.Context(0)
ctx.gp_set(0, X, Y, _lib.KERN_MATERN52, **synthetic.hypers(D))
ctx.program_set([[_lib.OP_EI, 0, 0, 0]], [-1.5])
ctx.set_variant(127)
Xc = synthetic.candidates(N, D)
ctx.eval(Xc[:65536], want_values=False)
vals, _, _, _ = ctx.eval(Xc)
t = ctx.timings()
c = vals.view(np.int64).ravel()[:148 * 16 * 8].reshape(148, 16, 8)
tot, gen, mma, bar, red = (c[:, :, i].astype(np.float64) for i in range(5))
print("kernel %.1f ms; per-warp cycles (mean over CTAs)" % t["hot_kernel_ms"])
for w in range(16):
    print("warp %2d (%s): total %.3e  gen %5.1f%%  mma %5.1f%%  barrier %5.1f%%  chunk-reduce %5.1f%%  other %5.1f%%" % (
        w, "gen-first" if ((w >> 2) & 1) == 0 else "mma-first", tot[:, w].mean(), 100 * gen[:, w].mean() / tot[:, w].mean(),
        100 * mma[:, w].mean() / tot[:, w].mean(), 100 * bar[:, w].mean() / tot[:, w].mean(), 100 * red[:, w].mean() / tot[:, w].mean(),
        100 * (tot[:, w] - gen[:, w] - mma[:, w] - bar[:, w] - red[:, w]).mean() / tot[:, w].mean()))
print("CTA total cycles: min %.3e max %.3e" % (tot[:, 0].min(), tot[:, 0].max()))
