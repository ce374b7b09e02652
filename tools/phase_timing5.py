"""Per-warp phase cycle counters of the profiling builds of the resident-panel kernel (variants 27 = v20, 28 = v21;
GPFO_BUILD_EXPERIMENTS=1).   python tools/phase_timing5.py [variant] [M]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflowopt_b200 import _lib, synthetic
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 27
M = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
D, N = 8, int(500000 * min(4.0, (2048.0 / M) ** 2))
X, Y = synthetic.training_set(M, D)
ctx = _lib.Context(0)
ctx.gp_set(0, X, Y, _lib.KERN_MATERN52, **synthetic.hypers(D))
ctx.program_set([[_lib.OP_EI, 0, 0, 0]], [-1.5])
ctx.set_variant(variant)
Xc = synthetic.candidates(N, D)
ctx.eval(Xc[:65536], want_values=False)
vals, _, _, _ = ctx.eval(Xc)
t = ctx.timings()
c = vals.view(np.int64).ravel()[:148 * 16 * 8].reshape(148, 16, 8).astype(np.float64)
tot = c[:, :, 0]
names = ["gen", "mma", "group-wait", "tile-prologue", "chunk-reduce", "epilogue+tile-barrier"]
print("variant %d M=%d N=%d: kernel %.1f ms; per-warp share of cycles (mean over CTAs)" % (variant, M, N, t["hot_kernel_ms"]))
for w in range(16):
    parts = [100 * c[:, w, 1 + i].mean() / tot[:, w].mean() for i in range(6)]
    print("warp %2d (%s): " % (w, "gen-first" if ((w >> 2) & 1) == 0 else "mma-first") +
          "  ".join("%s %5.1f%%" % (n, p) for n, p in zip(names, parts)) + "  other %5.1f%%" % (100 - sum(parts)))
print("CTA total cycles: min %.3e max %.3e" % (tot[:, 0].min(), tot[:, 0].max()))
