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
    for N in (1, 32, 1000):
        Xc = synthetic.candidates(N, D)
        tv = bench(lambda: acq.evaluate(Xc))
        tg = bench(lambda: acq.evaluate_with_gradients(Xc))
        ctx = acq._get_engine().ctx
        acq.evaluate_with_gradients(Xc)
        t = ctx.timings()
        print("M=%5d D=%2d N=%5d: evaluate %.3f ms, evaluate_with_gradients %.3f ms (device %.3f ms, %d launches)" % (
            M, D, N, tv, tg, t["eval_ms"], t["launches"]), flush=True)
