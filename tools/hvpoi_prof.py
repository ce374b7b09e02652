"""BASELINE configs[3] for ncu: HVPoI, 3 objectives, M = 1024 per GP, P ~ 200, one evaluation of N candidates."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflowopt_b200 import _lib, synthetic, pareto            # noqa: E402
from oracle import acq as oacq, gp as ogp                     # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
M, D, R = 1024, 6, 3
ctx = _lib.Context(0)
X = synthetic.candidates(M, D, seed=5)
F = synthetic.dtlz2_objectives(X, R, g_scale=0.4)
h = synthetic.hypers(D, 1e-3)
gps = [ogp.ScaledGPR(X, F[:, [j]], ogp.MATERN52, h["lengthscales"], h["variance"], h["noise_variance"]) for j in range(R)]
for j, g in enumerate(gps):
    ctx.gp_set(j, g.Xs, g.Ys, g.kind, g.lengthscales, g.variance, g.noise_variance, np.diag(g.input_transform.A),
               g.input_transform.b, np.diag(g.output_transform.A), g.output_transform.b, j)
ctx.gp_count_set(R)
Fm = np.hstack([ctx.predict(j, X, 1)[0] for j in range(R)])
p = pareto.Pareto(np.zeros((1, R)))
p.update(Fm)
hv = oacq.HVPoI(gps, np.asarray(p.front), np.asarray(p.bounds.lb), np.asarray(p.bounds.ub))
mlb, mub = pareto.merge_cells(hv.pf_ext(), hv.lb, hv.ub)
ctx.cells_set(hv.pf_ext(), mlb, mub)
ctx.program_set([[_lib.OP_HVPOI, 0, R, 0]], [])
Xc = synthetic.candidates(N, D)
vals, _, bv, bi = ctx.eval(Xc)
t = ctx.timings()
print("HVPoI N=%d P=%d C=%d: eval %.1f ms (contractions %.1f ms)" % (N, p.front.shape[0], mlb.shape[0], t["eval_ms"], t["hot_kernel_ms"]))
