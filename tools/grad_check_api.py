import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import gpflowopt_b200 as gpfo
from gpflowopt_b200 import synthetic
from gpflowopt_b200.acquisition import ExpectedImprovement
Ms, D = 4096, 8
X, Y = synthetic.training_set(Ms, D)
h = synthetic.hypers(D)
acq = ExpectedImprovement(gpfo.GPR(X, Y, gpfo.Matern52(D, variance=h["variance"], lengthscales=h["lengthscales"], ARD=True), noise_variance=h["noise_variance"]))
acq.optimize_restarts = 0
Xc = synthetic.candidates(125000, D)
acq.argmax(Xc[:65536]); acq.argmax(Xc)
ctx = acq._get_engine().ctx
for ng in (40000, 40000, 62500):
    acq.evaluate_with_gradients(Xc[:ng]); acq.evaluate_with_gradients(Xc[:ng])
    print(ng, ctx.timings())
