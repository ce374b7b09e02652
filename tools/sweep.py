"""BASELINE configs #2-#5 on one B200: throughput + oracle parity on a subset.  Writes gpurun_out/sweep.log"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflowopt_b200 import _lib, synthetic, pareto            # noqa: E402
from oracle import acq as oacq, gp as ogp                     # noqa: E402

PEAK = None


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


def upload(ctx, i, g, slot0):
    ctx.gp_set(i, g.Xs, g.Ys, g.kind, g.lengthscales, g.variance, g.noise_variance, np.diag(g.input_transform.A),
               g.input_transform.b, np.diag(g.output_transform.A), g.output_transform.b, slot0)


def model(M, D, kind=ogp.MATERN52, seed=1234, col=0, constraint=False):
    X, Y = synthetic.training_set(M, D, seed=seed, constraint=constraint)
    h = synthetic.hypers(D)
    return ogp.ScaledGPR(X, Y[:, [col]], kind, h["lengthscales"], h["variance"], h["noise_variance"])


def main():
    global PEAK
    ctx = _lib.Context(0)
    PEAK = ctx.measure_dmma_peak()
    print("FP64 DMMA peak %.2f TF/s" % PEAK, flush=True)
    # ---- config #2 / #5: EI, D = 8, M sweep, values and gradients ---------------------------------------------------
    for M in (512, 1024, 2048, 4096, 8192):
        g = model(M, 8)
        t0 = time.time()
        upload(ctx, 0, g, 0)
        ctx.gp_count_set(1)
        fm = ctx.timings()["factor_ms"]
        fmin = -1.5
        ctx.program_set([[_lib.OP_EI, 0, 0, 0]], [fmin])
        N = int(min(4e6, 1e6 * (2048.0 / M) ** 2))
        Xc = synthetic.candidates(N, 8)
        ctx.eval(Xc[:4096], want_values=False)
        vals, _, bv, bi = ctx.eval(Xc)
        t = ctx.timings()
        sub = np.random.RandomState(1).choice(N, 2000, replace=False)
        ref = oacq.Leaf('ei', g, fmin).value(Xc[sub])
        tf = N * M * (M + 1.0) / (t["hot_kernel_ms"] * 1e-3) / 1e12
        print("values  M=%5d N=%8d: factor %.1f ms, eval %.1f ms -> %.3e cand/s, %.2f TF/s (%.1f%% of peak), parity %.1e" % (
            M, N, fm, t["eval_ms"], N / (t["eval_ms"] * 1e-3), tf, 100 * tf / PEAK, rel(vals[sub], ref)), flush=True)
        Ng = max(40000, N // 8)
        ctx.eval(Xc[:2048], want_grads=True)
        ctx.eval(Xc[:Ng], want_grads=True)              # first call of this size: scratch allocation sits between the timing events
        vals, grads, _, _ = ctx.eval(Xc[:Ng], want_grads=True)
        t = ctx.timings()
        _, rg = oacq.Leaf('ei', g, fmin).value_and_grad(Xc[:200])
        tfg = Ng * 2.0 * M * (M + 1.0) / (t["eval_ms"] * 1e-3) / 1e12          # algorithmic F_grad
        print("grads   M=%5d N=%8d: eval %.1f ms -> %.3e cand/s, %.2f TF/s algorithmic 2M(M+1) (%.1f%% of peak, %s), grad parity %.1e" % (
            M, Ng, t["eval_ms"], Ng / (t["eval_ms"] * 1e-3), tfg, 100 * tfg / PEAK,
            "stored-A" if t["launches"] == 4 and Ng >= 37888 else "dense K^-1", rel(grads[:200], rg)), flush=True)
    # ---- config #5 upper end: N = 1e8 device-generated candidates (no N x D array anywhere), M = 512 ---------------
    M, N = 512, 10 ** 8
    g = model(M, 8)
    upload(ctx, 0, g, 0)
    ctx.gp_count_set(1)
    ctx.program_set([[_lib.OP_EI, 0, 0, 0]], [-1.5])
    lower, upper = np.zeros(8), np.ones(8)
    ctx.eval_random(4321, 0, 10 ** 6, lower, upper)
    _, bv, bi = ctx.eval_random(4321, 0, N, lower, upper)
    t = ctx.timings()
    xbest = _lib.random_rows(4321, bi, 1, lower, upper)
    first = 77000000
    vals, _, _ = ctx.eval_random(4321, first, 20000, lower, upper, want_values=True)
    rows = _lib.random_rows(4321, first, 20000, lower, upper)
    leaf = oacq.Leaf('ei', g, -1.5)
    print("N=1e8 device-generated, M=512: eval %.1f ms -> %.3e cand/s (%.2f TF/s, %.1f%% of peak); winner row %d regenerated on the host: "
          "oracle value rel. diff %.1e; parity on regenerated rows %d..%d: %.1e" % (
              t["eval_ms"], N / (t["eval_ms"] * 1e-3), N * M * (M + 1.0) / (t["eval_ms"] * 1e-3) / 1e12,
              100 * N * M * (M + 1.0) / (t["eval_ms"] * 1e-3) / 1e12 / PEAK, bi,
              abs(leaf.value(xbest)[0, 0] - bv) / abs(bv), first, first + 20000, rel(vals, leaf.value(rows))), flush=True)
    # ---- config #3: EI x PoF, two GPs, M = 4096, D = 16 (one GPU's shard of the 1e7 / 8-GPU job) -----------------
    M, D, N = 4096, 16, 1250000
    g1, g2 = model(M, D, constraint=True, col=0), model(M, D, constraint=True, col=1)
    upload(ctx, 0, g1, 0)
    upload(ctx, 1, g2, 1)
    ctx.gp_count_set(2)
    fmin = -1.0
    ctx.program_set([[_lib.OP_EI, 0, 0, 0], [_lib.OP_POF, 1, 0, 1], [_lib.OP_PROD, 2, 0, 0]], [fmin, 0.0])
    Xc = synthetic.candidates(N, D)
    ctx.eval(Xc[:4096], want_values=False)
    vals, _, bv, bi = ctx.eval(Xc)
    t = ctx.timings()
    sub = np.random.RandomState(2).choice(N, 1000, replace=False)
    ref = oacq.Agg('prod', [oacq.Leaf('ei', g1, fmin), oacq.Leaf('pof', g2, 0.0)]).value(Xc[sub])
    tf = 2.0 * N * M * (M + 1.0) / (t["hot_kernel_ms"] * 1e-3) / 1e12
    print("config3 EIxPoF M=%d D=%d N=%d (1/8 of 1e7): eval %.1f ms -> %.3e cand/s, %.2f TF/s (%.1f%%), parity %.1e, argmax==argmax(vals) %s" % (
        M, D, N, t["eval_ms"], N / (t["eval_ms"] * 1e-3), tf, 100 * tf / PEAK, rel(vals[sub], ref), bi == int(np.argmax(vals))), flush=True)
    # ---- config #4: HVPoI, 3 objectives, M = 1024 per GP, P ~ 200 ---------------------------------------------------
    M, D, R, N = 1024, 4, 3, 1000000
    X = synthetic.candidates(M, D, seed=5)
    F = synthetic.dtlz2_objectives(X, R)
    h = synthetic.hypers(D, 1e-3)
    gps = [ogp.ScaledGPR(X, F[:, [j]], ogp.MATERN52, h["lengthscales"], h["variance"], h["noise_variance"]) for j in range(R)]
    for j, g in enumerate(gps):
        upload(ctx, j, g, j)
    ctx.gp_count_set(R)
    Fm = np.hstack([ctx.predict(j, X, 1)[0] for j in range(R)])
    t0 = time.time()
    p = pareto.Pareto(np.zeros((1, R)))
    p.update(Fm)
    t_pareto = time.time() - t0
    hv = oacq.HVPoI(gps, np.asarray(p.front), np.asarray(p.bounds.lb), np.asarray(p.bounds.ub))
    t0 = time.time()
    mlb, mub = pareto.merge_cells(hv.pf_ext(), hv.lb, hv.ub)        # what HVProbabilityOfImprovement uploads
    t_merge = time.time() - t0
    ctx.cells_set(hv.pf_ext(), mlb, mub)
    ctx.program_set([[_lib.OP_HVPOI, 0, R, 0]], [])
    Xc = synthetic.candidates(N, D)
    ctx.eval(Xc[:4096], want_values=False)
    vals, _, bv, bi = ctx.eval(Xc)
    t = ctx.timings()
    sub = np.random.RandomState(3).choice(N, 300, replace=False)
    ref = hv.value(Xc[sub])
    C = mlb.shape[0]
    print("config4 cells: reference decomposition %d -> coarsened %d (%.3f s)" % (hv.lb.shape[0], C, t_merge))
    cell_ms = t["eval_ms"] - t["hot_kernel_ms"]
    print("config4 HVPoI R=3 M=%d N=%d P=%d C=%d (host cell decomposition %.1f s): eval %.1f ms (GP contractions %.1f, cells %.1f) -> %.3e cand/s; "
          "cell kernel %.2e cell*cand/s; parity %.1e" % (M, N, p.front.shape[0], C, t_pareto, t["eval_ms"], t["hot_kernel_ms"], cell_ms,
                                                       N / (t["eval_ms"] * 1e-3), N * C / (cell_ms * 1e-3), rel(vals[sub], ref)), flush=True)
    vals, grads, _, _ = ctx.eval(Xc[:20000], want_grads=True)
    t = ctx.timings()
    rv, rg = hv.value_and_grad(Xc[:64])
    print("config4 HVPoI gradients N=20000: eval %.1f ms, grad parity %.1e" % (t["eval_ms"], rel(grads[:64], rg)), flush=True)


if __name__ == "__main__":
    main()
