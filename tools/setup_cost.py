"""Per-BO-iteration setup cost: factorisation + packing (gpfo_gp_set) and the lazily built gradient-pass state."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflowopt_b200 import _lib, synthetic           # noqa: E402

ctx = _lib.Context(0)
for M in (512, 2048, 4096, 8192):
    D = 8
    X, Y = synthetic.training_set(M, D)
    ctx.gp_set(0, X, Y, _lib.KERN_MATERN52, **synthetic.hypers(D))          # warm allocations
    ctx.program_set([[_lib.OP_EI, 0, 0, 0]], [-1.5])
    Xs, Xl = synthetic.candidates(1, D), synthetic.candidates(40000, D)
    ctx.eval(Xs, want_grads=True); ctx.eval(Xl, want_grads=True)
    t0 = time.perf_counter(); ctx.gp_set(0, X, Y, _lib.KERN_MATERN52, **synthetic.hypers(D)); t_set = time.perf_counter() - t0
    fm = ctx.timings()["factor_ms"]
    t0 = time.perf_counter(); ctx.eval(Xs, want_grads=True); t_g1 = time.perf_counter() - t0      # builds K^-1 + alpha + 2 packs
    t0 = time.perf_counter(); ctx.eval(Xs, want_grads=True); t_g2 = time.perf_counter() - t0
    t0 = time.perf_counter(); ctx.eval(Xl, want_grads=True); t_l1 = time.perf_counter() - t0      # builds the reversed transposed stream
    t0 = time.perf_counter(); ctx.eval(Xl, want_grads=True); t_l2 = time.perf_counter() - t0
    print("M=%5d: gp_set %.1f ms (device factor+pack %.1f ms); first N=1 gradient call %.1f ms vs %.2f ms after; "
          "first large gradient batch %.1f ms vs %.1f ms after" % (M, t_set * 1e3, fm, t_g1 * 1e3, t_g2 * 1e3, t_l1 * 1e3, t_l2 * 1e3), flush=True)
