"""Latency of small calls through the public API (the L-BFGS-B polish steps and BASELINE configs[0])."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpflowopt_b200 as gpfo
from gpflowopt_b200 import synthetic
from gpflowopt_b200.acquisition import ExpectedImprovement


def bench(f, reps=50):
    f(); f()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    return (time.perf_counter() - t0) / reps * 1e3


for M, D in ((20, 2), (512, 8), (2048, 8), (4096, 16)):
    X, Y = synthetic.training_set(M, D)
    h = synthetic.hypers(D)
    kern = gpfo.RBF(D, variance=h["variance"], lengthscales=h["lengthscales"], ARD=True) if M == 20 else \
        gpfo.Matern52(D, variance=h["variance"], lengthscales=h["lengthscales"], ARD=True)
    acq = ExpectedImprovement(gpfo.GPR(X, Y, kern, noise_variance=h["noise_variance"]))
    for N in (1, 2, 4, 8, 32, 1000):
        Xc = synthetic.candidates(N, D)
        tv = bench(lambda: acq.evaluate(Xc))
        tg = bench(lambda: acq.evaluate_with_gradients(Xc))
        ctx = acq._get_engine().ctx
        acq.evaluate_with_gradients(Xc)
        t = ctx.timings()
        print("M=%5d D=%2d N=%5d: evaluate %.3f ms, evaluate_with_gradients %.3f ms (device %.3f ms, %d launches)" % (
            M, D, N, tv, tg, t["eval_ms"], t["launches"]), flush=True)

# ---- EI x PoF (two GPs, BASELINE configs[2] model) at M = 4096, D = 16: the polish call of a constrained BO iteration ----
from gpflowopt_b200.acquisition import ProbabilityOfFeasibility                       # noqa: E402
M, D = 4096, 16
X, Y = synthetic.training_set(M, D, constraint=True)
h = synthetic.hypers(D)
mk = lambda y: gpfo.GPR(X, y, gpfo.Matern52(D, variance=h["variance"], lengthscales=h["lengthscales"], ARD=True), noise_variance=h["noise_variance"])
joint = ExpectedImprovement(mk(Y[:, [0]])) * ProbabilityOfFeasibility(mk(Y[:, [1]]))
for N in (1, 8):
    Xc = synthetic.candidates(N, D)
    tv = bench(lambda: joint.evaluate(Xc))
    tg = bench(lambda: joint.evaluate_with_gradients(Xc))
    joint.evaluate_with_gradients(Xc)
    t = joint._get_engine().ctx.timings()
    print("EI x PoF M=%5d D=%2d N=%5d: evaluate %.3f ms, evaluate_with_gradients %.3f ms (device %.3f ms, %d launches)" % (
        M, D, N, tv, tg, t["eval_ms"], t["launches"]), flush=True)

# ---- one BO inner optimisation (BASELINE configs[1] model): MC stage + polish, reference-style vs batched multi-start ----
from gpflowopt_b200.bo import _InverseAcquisition                                   # noqa: E402
from gpflowopt_b200.domain import ContinuousParameter                               # noqa: E402
from gpflowopt_b200.optim import MCOptimizer, MultiStartOptimizer, SciPyOptimizer, StagedOptimizer   # noqa: E402

M, D, NMC = 2048, 8, 10 ** 6
X, Y = synthetic.training_set(M, D)
h = synthetic.hypers(D)
acq = ExpectedImprovement(gpfo.GPR(X, Y, gpfo.Matern52(D, variance=h["variance"], lengthscales=h["lengthscales"], ARD=True),
                                   noise_variance=h["noise_variance"]))
dom = np.sum([ContinuousParameter("x%d" % i, 0.0, 1.0) for i in range(D)])
staged = StagedOptimizer([MCOptimizer(dom, NMC), SciPyOptimizer(dom)])
staged_dev = StagedOptimizer([MCOptimizer(dom, NMC, device_rng=True), SciPyOptimizer(dom)])
multi = MultiStartOptimizer(dom, NMC, nstarts=8)
multi_dev = MultiStartOptimizer(dom, NMC, nstarts=8, device_rng=True)
for name, opt in (("staged MC(1e6) + L-BFGS-B (reference pipeline)", staged), ("same, device-generated candidates", staged_dev),
                  ("MC(1e6) + 8-start lock-step L-BFGS-B", multi), ("same, device-generated candidates", multi_dev)):
    best = None
    for rep in range(3):
        np.random.seed(100 + rep)
        t0 = time.perf_counter()
        r = opt.optimize(_InverseAcquisition(acq, opt.gradient_enabled()))
        dt = (time.perf_counter() - t0) * 1e3
        best = dt if best is None else min(best, dt)
    extra = " (%d batched gradient calls for %d rows)" % (r.nbatches, r.polish_rows) if hasattr(r, "nbatches") else ""
    print("%-50s: %.1f ms per inner optimisation, acq = %.6g, nfev = %d%s" % (name, best, -float(np.ravel(r.fun)[0]), r.nfev, extra),
          flush=True)
