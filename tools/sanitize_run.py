"""Small end-to-end pass for compute-sanitizer: every kernel family once (factor, contraction variants, gradient pass, HVPoI)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflowopt_b200 import _lib, synthetic, pareto
ctx = _lib.Context(0)
M, D = 300, 5
X, Y = synthetic.training_set(M, D)
ctx.gp_set(0, X, Y, _lib.KERN_MATERN52, **synthetic.hypers(D))
ctx.program_set([[_lib.OP_EI, 0, 0, 0]], [-1.0])
Xc = synthetic.candidates(int(sys.argv[1]) if len(sys.argv) > 1 else 700, D)
for v in (4, 5, 6, 11, 12, 10):
    ctx.set_variant(v)
    vals, _, bv, bi = ctx.eval(Xc)
    print("variant", v, float(vals.sum()), bi)
ctx.set_variant(0)
vals, grads, _, _ = ctx.eval(Xc[:200], want_grads=True)
print("grad", float(grads.sum()))
R = 3
F = synthetic.dtlz2_objectives(X[:, :4], R)
for j in range(R):
    ctx.gp_set(j, X, F[:, [j]], _lib.KERN_RBF, slot0=j, **synthetic.hypers(D, 1e-3))
ctx.gp_count_set(R)
p = pareto.Pareto(np.zeros((1, R)))
p.update(F)
front = np.asarray(p.front)
ref = front.max(0, keepdims=True) + 2 * (front.max(0, keepdims=True) - front.min(0, keepdims=True)) / front.shape[0]
pf_ext = np.vstack((-np.inf * np.ones((1, R)), front, ref))
mlb, mub = pareto.merge_cells(pf_ext, p.bounds.lb, p.bounds.ub)
ctx.cells_set(pf_ext, mlb, mub)
ctx.program_set([[_lib.OP_HVPOI, 0, R, 0]], [])
vals, grads, _, _ = ctx.eval(Xc[:100], want_grads=True)
print("hvpoi", float(vals.sum()), float(grads.sum()))
