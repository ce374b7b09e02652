#!/usr/bin/env python
"""bench.py -- acquisition candidate evaluations / second on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch of synthetic candidates: BASELINE.json configs[1]
(EI, GPR Matern52-ARD, M=2048, D=8, float64, 1e6 random candidates per GPU per step).  Weak scaling: every rank
scores its own 1e6-row shard and the ranks exchange one 16-byte (value, index) pair per step (NCCL all-gather).

Printed JSON (rank 0, one line):
  value      whole-job candidates/s, inputs resident in HBM, device time (CUDA events on the library's stream)
  e2e        the same metric through the public Python API (Acquisition.argmax) with pinned HOST candidates:
             H2D of the candidates and D2H of the (value, index) result inside the timed region
  roofline   the DMMA contraction kernel vs this GPU's FP64 tensor peak measured live (register-only mma.sync loop)
  cpu_baseline  the CPU oracle (NumPy/SciPy restatement of the reference path) on the host cores, bounded sample
  configs    sub-records for the other BASELINE configs under the same clock: configs[2] (EI x PoF, M=4096, D=16, 1e7
             candidates STRONG-scaled over the ranks, host-candidate and device-RNG routes, device time and wall),
             configs[3] (HVPoI: GP contractions and cell kernel), configs[4] (values + gradients over M = 512...8192)
  bars       library / CPU sanity bars of SURVEY 8(d): torch-fp64 `Linv @ K(X,X*)` on this GPU, best-of-10 FP64 GEMM 8192^3
             (cuBLAS through torch.matmul) beside the DMMA issue peak, 1-thread CPU, per-call-Cholesky CPU, pageable e2e
(--no-extras prints the headline only.)
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M, D = 2048, 8
N_PER_GPU = 1000000
METRIC = "acq candidate evals/sec (M=2048,D=8,fp64)"
UNIT = "candidates/s"
# The forward kernels a large batch can select (gpfo_timings.variant): name, and dram__bytes_read.sum + dram__bytes_write.sum
# of ONE launch at this workload (N = 1e6, M = 2048) from the ncu capture named beside it -- DRAM bytes cannot be measured
# outside a profiler, so the figure is a cited measurement, never a constant invented here.
KERNELS = {
    20: {"kernel": "contract5_kernel<8,2,2> (FP64 DMMA.8x8x4 contraction + fused EI epilogue + block argmax; 16-candidate tiles, "
                   "1024-row chunks, K(X*,X) generated once per tile into a resident shared-memory panel, no global scratch)",
         "traffic": 81803264 + 3297792, "traffic_source": "profiles/r02_final_contract5_v20_dram_N1e6.csv (2026-10-20, final build; N = 1e6, M = 2048)"},
    21: {"kernel": "contract5_kernel<4,4,2> (as variant 20 with 32-candidate tiles and 512-row chunks)",
         "traffic": None, "traffic_source": None},
    12: {"kernel": "contract2_kernel<16,2,8,5,*,8> (FP64 DMMA.8x8x4 contraction + fused EI epilogue + block argmax; 64-candidate "
                   "tiles, 256-row chunks, K(X*,X) generated once per tile and parked in a per-CTA global scratch)",
         "traffic": 27688939000 + 16286433000, "traffic_source": "profiles/r01_final_contract2_v12.metrics.csv (2026-10-19)"},
}


def _workload_text(n):
    return "EI on %d random candidates per GPU per step, D=8, GPR Matern52-ARD M=2048, float64 (BASELINE configs[1])" % n


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super(ClockSampler, self).__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._halt = index, [], set(), None, threading.Event()

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                     "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
            while not self._halt.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
                self._halt.wait(0.1)
        except Exception as e:                         # pragma: no cover
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def finish(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def host_threads():
    """Threads the CPU arm really uses: all host cores (torchrun exports OMP_NUM_THREADS=1, so the BLAS pool is
    re-sized at run time through threadpoolctl)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:                                 # pragma: no cover
        n = os.cpu_count() or 1
    return max(1, n)


def cpu_oracle_throughput(n_sample):
    """Times the CPU oracle (op-for-op NumPy/SciPy restatement; OpenBLAS threads = all host cores) on a bounded
    sample of the same workload; candidates in chunks of 4096 rows (what the literal reference graph can hold)."""
    from threadpoolctl import threadpool_limits
    with threadpool_limits(limits=host_threads()):
        return _cpu_oracle_throughput(n_sample)


def _cpu_oracle_throughput(n_sample):
    from gpflowopt_b200 import synthetic
    from oracle import acq as oacq, gp as ogp
    X, Y = synthetic.training_set(M, D)
    h = synthetic.hypers(D)
    g = ogp.ScaledGPR(X, Y, ogp.MATERN52, h["lengthscales"], h["variance"], h["noise_variance"])
    g.factor()                                           # hoisted (the reference re-factorises every call)
    ei = oacq.Leaf('ei', g, -1.5)
    Xc = synthetic.candidates(n_sample, D)
    t0 = time.perf_counter()
    best = -np.inf
    for s in range(0, n_sample, 4096):
        best = max(best, float(np.max(ei.value(Xc[s:s + 4096]))))
    dt = time.perf_counter() - t0
    return n_sample / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    n_sample = 20000
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_oracle_throughput(4096)
    times = []
    for _ in range(args.steps):
        v, dt = cpu_oracle_throughput(n_sample)
        times.append(dt)
    dt = float(np.mean(times))
    value = n_sample / dt
    sample = "%d candidates per step in 4096-row chunks, Cholesky hoisted, NumPy/SciPy + OpenBLAS" % n_sample
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": _workload_text(N_PER_GPU), "M": M, "D": D, "sample_candidates_per_step": n_sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "GPflowOpt 0.1.1 needs GPflow 0.5.0 + TensorFlow 1.x (absent, not installable): the reference arm is "
                "the CPU oracle port of its path (oracle/), all host threads via OpenBLAS"})


# ---------------------------------------------------------------------------------------------------------------------
# Sub-records: the other BASELINE configs and the library / CPU bars of SURVEY 8(d), all under the same driver clock
# ---------------------------------------------------------------------------------------------------------------------
def cpu_side_bars():
    """1-thread CPU rate and the "reference-faithful" rate (Cholesky re-run inside every 4096-row call, as every
    session.run of the reference does: SURVEY Appendix B), both with the oracle port on a bounded sample."""
    from threadpoolctl import threadpool_limits
    from gpflowopt_b200 import synthetic
    from oracle import acq as oacq, gp as ogp
    out = {}
    with threadpool_limits(limits=1):
        v1, dt1 = _cpu_oracle_throughput(4096)
    out["cpu_1_thread"] = {"value": v1, "unit": UNIT, "cores": 1, "sample": "4096 candidates (%.1f s), Cholesky hoisted" % dt1}
    with threadpool_limits(limits=host_threads()):
        X, Y = synthetic.training_set(M, D)
        h = synthetic.hypers(D)
        Xc = synthetic.candidates(2 * 4096, D)
        t0 = time.perf_counter()
        for s in range(0, Xc.shape[0], 4096):
            g = ogp.ScaledGPR(X, Y, ogp.MATERN52, h["lengthscales"], h["variance"], h["noise_variance"])   # fresh: K, chol per call
            oacq.Leaf('ei', g, -1.5).value(Xc[s:s + 4096])
        dt = time.perf_counter() - t0
    out["cpu_per_call_cholesky"] = {"value": Xc.shape[0] / dt, "unit": UNIT, "cores": host_threads(),
                                    "sample": "2 calls of 4096 candidates (%.1f s), K + Cholesky rebuilt inside each call "
                                              "as the reference's session.run does" % dt}
    return out


def gpu_library_bars(torch, ctx_peak):
    """(1) torch-fp64 restatement of the same scoring on this GPU with library kernels only (cdist-style r^2, Matern52,
    `Linv @ Kx` through cuBLAS DGEMM, reductions): the "library bar".  (2) best-of-10 FP64 GEMM 8192^3 through torch.matmul
    (cuBLAS): the library's FP64 tensor throughput beside the register-only DMMA issue peak."""
    from gpflowopt_b200 import synthetic
    dev = torch.device("cuda", torch.cuda.current_device())
    X, Y = synthetic.training_set(M, D)
    h = synthetic.hypers(D)
    Xs = torch.from_numpy(X / h["lengthscales"]).to(dev)
    r2 = (Xs * Xs).sum(1)[:, None] + (Xs * Xs).sum(1)[None, :] - 2.0 * Xs @ Xs.T
    r = torch.sqrt(torch.clamp(r2, min=0.0) + 1e-12)
    s5 = 5.0 ** 0.5
    K = h["variance"] * (1.0 + s5 * r + (5.0 / 3.0) * r * r) * torch.exp(-s5 * r) + h["noise_variance"] * torch.eye(M, dtype=torch.float64, device=dev)
    L = torch.linalg.cholesky(K)
    Linv = torch.linalg.solve_triangular(L, torch.eye(M, dtype=torch.float64, device=dev), upper=False)
    V = Linv @ torch.from_numpy(Y).to(dev)
    ell = torch.from_numpy(h["lengthscales"]).to(dev)
    n_lib, chunk = 131072, 16384
    Xc = torch.from_numpy(synthetic.candidates(n_lib, D)).to(dev)
    xn = (Xs * Xs).sum(1)[:, None]
    fmin = -1.5

    def score(xc):
        xs = xc / ell
        r2c = xn + (xs * xs).sum(1)[None, :] - 2.0 * Xs @ xs.T                   # M x n
        rc = torch.sqrt(r2c + 1e-12)
        Kx = h["variance"] * (1.0 + s5 * rc + (5.0 / 3.0) * rc * rc) * torch.exp(-s5 * rc)
        A = Linv @ Kx                                                          # the dense N x M x M term (cuBLAS DGEMM)
        mean = (A.T @ V)[:, 0]
        var = torch.clamp(h["variance"] - (A * A).sum(0), min=1e-6)
        sd = torch.sqrt(var)
        z = (fmin - mean) / sd
        ei = (fmin - mean) * 0.5 * (1.0 + torch.erf(z / 2.0 ** 0.5)) + sd * torch.exp(-0.5 * z * z) / (2.0 * np.pi) ** 0.5
        return ei.max()

    score(Xc[:chunk])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(0, n_lib, chunk):
        score(Xc[s:s + chunk])
    e1.record()
    torch.cuda.synchronize()
    lib_ms = e0.elapsed_time(e1)
    a = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    b = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    best = 1e30
    for _ in range(10):
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    dgemm = 2.0 * 8192.0 ** 3 / (best * 1e-3) / 1e12
    del a, b
    return {
        "torch_fp64_library": {"value": n_lib / (lib_ms * 1e-3), "unit": UNIT,
                               "sample": "%d candidates in %d-row chunks, same M/D/kernel; Kx materialised, Linv @ Kx = dense "
                                         "cuBLAS DGEMM (2 M^2 flop/candidate), torch elementwise ops" % (n_lib, chunk),
                               "tflops_dense": 2.0 * M * M * n_lib / (lib_ms * 1e-3) / 1e12},
        "fp64_gemm_8192_best_of_10": {"value": dgemm, "unit": "TFLOP/s", "how": "torch.matmul float64 (cuBLAS), CUDA events"},
        "dmma_issue_peak": {"value": ctx_peak, "unit": "TFLOP/s", "how": "gpfo_measure_dmma_peak: register-only mma.sync m8n8k4 loop"},
    }


def sub_config3_constrained(torch, td, gpfo, world, rank, barrier):
    """BASELINE configs[2]: EI x PoF (2 GPs, M = 4096, D = 16) over 1e7 candidates, STRONG-scaled over the ranks, through the
    public optimiser API.  Two candidate routes: the reference's (rank 0 draws np.random rows, pipelined scatter) and the
    device generator.  Wall clock around Optimizer.optimize (max over ranks) and summed device time of the scoring calls."""
    from gpflowopt_b200 import synthetic
    from gpflowopt_b200.acquisition import ExpectedImprovement, ProbabilityOfFeasibility
    from gpflowopt_b200.bo import _InverseAcquisition
    from gpflowopt_b200.domain import UnitCube
    from gpflowopt_b200.optim import MCOptimizer
    M3, D3, N3 = 4096, 16, 10 ** 7
    X, Y = synthetic.training_set(M3, D3, constraint=True)
    h = synthetic.hypers(D3)

    def mk(c):
        return gpfo.GPR(X, Y[:, [c]], gpfo.Matern52(D3, variance=h["variance"], lengthscales=h["lengthscales"], ARD=True),
                        noise_variance=h["noise_variance"])
    t0 = time.perf_counter()
    ei, pof = ExpectedImprovement(mk(0)), ProbabilityOfFeasibility(mk(1))
    ei.optimize_restarts = pof.optimize_restarts = 0
    joint = ei * pof
    joint.evaluate(np.full((1, D3), 0.5))                               # setup: two factorisations, feasibility, fmin, program
    setup_s = time.perf_counter() - t0
    ctx = joint._get_engine().ctx
    dom = UnitCube(D3)
    rec = {"workload": "EI x PoF, 2 GPs, Matern52-ARD M=4096, D=16, 1e7 candidates over %d GPU(s) (BASELINE configs[2])" % world,
           "scaling": "strong", "n_gpus": world, "candidates_total": N3, "setup_s": setup_s,
           "flop_per_candidate": 2 * M3 * (M3 + 1)}
    winners = {}
    for key, opt in (("host_candidates", MCOptimizer(dom, N3)), ("device_rng", MCOptimizer(dom, N3, device_rng=True))):
        # warm-up on a small candidate set (as the W warm-up steps of the headline): the first scatter of a process pays NCCL's
        # lazy point-to-point connection set-up (about 1.1 s on 8 GPUs, once per process, not per BO iteration)
        np.random.seed(1)
        MCOptimizer(dom, 65536, device_rng=(key == "device_rng")).optimize(_InverseAcquisition(joint, False))
        np.random.seed(4321)
        ctx.reset_cum()
        barrier()
        t0 = time.perf_counter()
        res = opt.optimize(_InverseAcquisition(joint, False))
        barrier()
        wall = time.perf_counter() - t0
        st = torch.tensor([wall, ctx.cum["eval_ms"], ctx.cum["hot_kernel_ms"]], dtype=torch.float64, device="cuda")
        if world > 1:
            td.all_reduce(st, op=td.ReduceOp.MAX)
        wall, dev_ms, hot_ms = [float(v) for v in st.cpu()]
        winners[key] = (res.x.copy(), float(np.ravel(res.fun)[0]))
        rec[key] = {"wall_s": wall, "value_wall": N3 / wall, "device_ms": dev_ms, "value_device": N3 / (dev_ms * 1e-3),
                    "unit": UNIT, "tflops_per_gpu": rec["flop_per_candidate"] * (N3 / world) / (hot_ms * 1e-3) / 1e12,
                    "scoring_calls_per_rank": ctx.cum["calls"], "best_acq": -winners[key][1]}
    rec["host_candidates"]["route"] = ("rank 0 draws np.random.rand rows in order (the reference's single stream, design.py:83-94), "
                                       "8 pipelined scatter rounds; every rank scores its chunk while the next is drawn")
    rec["device_rng"]["route"] = "counter-based generator inside the scoring kernel, every rank its own rows (opt-in, different stream)"
    joint._get_engine().close()
    return rec


def sub_config4_hvpoi(gpfo):
    """BASELINE configs[3]: HVProbabilityOfImprovement, 3 objectives, M = 1024 per GP, D = 6, 1e6 candidates, front ~200."""
    from gpflowopt_b200 import synthetic
    from gpflowopt_b200.acquisition import HVProbabilityOfImprovement
    M4, D4, R4, N4 = 1024, 6, 3, 1000000
    X = synthetic.candidates(M4, D4, seed=5)
    F = synthetic.dtlz2_objectives(X, R4, g_scale=0.4)
    h = synthetic.hypers(D4, 1e-3)
    models = [gpfo.GPR(X, F[:, [j]], gpfo.Matern52(D4, variance=h["variance"], lengthscales=h["lengthscales"], ARD=True),
                       noise_variance=h["noise_variance"]) for j in range(R4)]
    t0 = time.perf_counter()
    hv = HVProbabilityOfImprovement(models)
    hv.optimize_restarts = 0
    Xc = synthetic.candidates(N4, D4)
    hv.evaluate(Xc[:256])                                               # setup: 3 factorisations, front, cell decomposition
    setup_s = time.perf_counter() - t0
    ctx = hv._get_engine().ctx
    hv.argmax(Xc)
    best = None
    for _ in range(2):
        v, i = hv.argmax(Xc)
        t = ctx.timings()
        if best is None or t["eval_ms"] < best["eval_ms"]:
            best = t
    P = int(hv.pareto.front.shape[0]) if hasattr(hv, "pareto") else None
    rec = {"workload": "HVPoI, 3 objectives, Matern52-ARD M=1024 per GP, D=6, 1e6 candidates (BASELINE configs[3])",
           "pareto_front": P, "setup_s": setup_s, "eval_ms": best["eval_ms"], "gp_contractions_ms": best["hot_kernel_ms"],
           "cell_kernel_ms": best["eval_ms"] - best["hot_kernel_ms"], "value": N4 / (best["eval_ms"] * 1e-3), "unit": UNIT,
           "launches": int(best["launches"]), "best": {"value": v, "index": i}}
    hv._get_engine().close()
    return rec


def sub_config5_sweep(gpfo, peak):
    """BASELINE configs[4] on one GPU: M = 512 ... 8192, values (argmax) and evaluate_with_gradients, device time."""
    from gpflowopt_b200 import synthetic
    from gpflowopt_b200.acquisition import ExpectedImprovement
    rows = []
    for Ms in (512, 1024, 2048, 4096, 8192):
        X, Y = synthetic.training_set(Ms, D)
        h = synthetic.hypers(D)
        acq = ExpectedImprovement(gpfo.GPR(X, Y, gpfo.Matern52(D, variance=h["variance"], lengthscales=h["lengthscales"], ARD=True),
                                           noise_variance=h["noise_variance"]))
        acq.optimize_restarts = 0
        n = max(40000, int(min(2e6, 5e5 * (2048.0 / Ms) ** 2)))     # >= 256 candidates per SM: the large-batch kernels
        ng = max(40000, n // 4)
        Xc = synthetic.candidates(n, D)
        acq.argmax(Xc[:65536])
        ctx = acq._get_engine().ctx
        t = tg = None
        for _ in range(3):                               # best of three (the first call after host-side set-up sees ramping clocks)
            acq.argmax(Xc)
            if t is None or ctx.timings()["eval_ms"] < t["eval_ms"]:
                t = ctx.timings()
        tf = n * Ms * (Ms + 1.0) / (t["hot_kernel_ms"] * 1e-3) / 1e12
        acq.evaluate_with_gradients(Xc[:ng])            # builds the gradient-pass state of the model (lazily, once)
        g_calls = []
        for _ in range(3):
            acq.evaluate_with_gradients(Xc[:ng])
            g_calls.append(round(ctx.timings()["eval_ms"], 3))
            if tg is None or ctx.timings()["eval_ms"] < tg["eval_ms"]:
                tg = ctx.timings()
        tfg = ng * 2.0 * Ms * (Ms + 1.0) / (tg["eval_ms"] * 1e-3) / 1e12
        rows.append({"M": Ms, "values": {"candidates": n, "eval_ms": t["eval_ms"], "value": n / (t["eval_ms"] * 1e-3),
                                         "tflops": tf, "frac_of_dmma_peak": tf / peak, "kernel_variant": int(t["variant"])},
                     "gradients": {"candidates": ng, "eval_ms": tg["eval_ms"], "eval_ms_calls": g_calls, "launches": int(tg["launches"]),
                                   "value": ng / (tg["eval_ms"] * 1e-3),
                                   "tflops_algorithmic_2M2": tfg, "frac_of_dmma_peak": tfg / peak}})
        acq._get_engine().close()
    return {"workload": "EI, Matern52-ARD, D=8, M sweep, values and evaluate_with_gradients on one GPU (BASELINE configs[4])",
            "unit": UNIT, "timing": "device time (gpfo_timings.eval_ms), best of 3 calls", "rows": rows}


def sub_config1_branin(gpfo):
    """BASELINE configs[0]: ExpectedImprovement on 2-D Branin, GPR RBF-ARD with 20 training points, MCOptimizer(domain, 1000)
    (reference optim.py:152-164): wall clock of one whole optimize() call through the public API -- 1000 candidates drawn with
    np.random on the host, scored on the GPU, arg-min returned."""
    import numpy as np
    from gpflowopt_b200.acquisition import ExpectedImprovement
    from gpflowopt_b200.bo import _InverseAcquisition
    from gpflowopt_b200.domain import ContinuousParameter
    from gpflowopt_b200.optim import MCOptimizer
    domain = ContinuousParameter('x1', -5, 10) + ContinuousParameter('x2', 0, 15)
    rs = np.random.RandomState(7)
    X = domain.lower + rs.rand(20, 2) * (domain.upper - domain.lower)
    x1, x2 = X[:, 0], X[:, 1]
    Y = ((x2 - 5.1 / (4 * np.pi ** 2) * x1 ** 2 + 5 / np.pi * x1 - 6) ** 2 + 10 * (1 - 1 / (8 * np.pi)) * np.cos(x1) + 10)[:, None]
    acq = ExpectedImprovement(gpfo.GPR(X, Y, gpfo.RBF(2, variance=1.0, lengthscales=np.array([0.3, 0.4]), ARD=True), noise_variance=1e-3))
    acq.optimize_restarts = 0
    acq.enable_scaling(domain)
    opt, obj = MCOptimizer(domain, 1000), _InverseAcquisition(acq, False)
    np.random.seed(11)
    for _ in range(5):
        res = opt.optimize(obj)
    reps = 200
    t0 = time.perf_counter()
    for _ in range(reps):
        res = opt.optimize(obj)
    ms = (time.perf_counter() - t0) / reps * 1e3
    rec = {"workload": "EI, GPR RBF-ARD M=20, 2-D Branin, MCOptimizer(domain, 1000).optimize (BASELINE configs[0])",
           "optimize_ms": ms, "value": 1000.0 / (ms * 1e-3), "unit": UNIT, "timing": "wall clock per optimize() call, mean of %d" % reps,
           "device_ms": acq._get_engine().ctx.timings()["eval_ms"], "fun": float(res.fun[0, 0])}
    acq._get_engine().close()
    return rec


def sub_polish_latency(gpfo):
    """The N = 1 calls an L-BFGS-B polish makes (reference optim.py:214-224 through acquisition.py:256-257): wall clock of
    Acquisition.evaluate / evaluate_with_gradients of one candidate through the public API, host arrays in and out."""
    from gpflowopt_b200 import synthetic
    from gpflowopt_b200.acquisition import ExpectedImprovement, ProbabilityOfFeasibility

    def wall_ms(f, reps=200):
        for _ in range(5):
            f()
        t0 = time.perf_counter()
        for _ in range(reps):
            f()
        return (time.perf_counter() - t0) / reps * 1e3

    rows = []
    for name, Ms, Ds, joint in (("EI", 2048, 8, False), ("EI x PoF", 4096, 16, True)):
        X, Y = synthetic.training_set(Ms, Ds, constraint=joint)
        h = synthetic.hypers(Ds)
        mk = lambda y: gpfo.GPR(X, y, gpfo.Matern52(Ds, variance=h["variance"], lengthscales=h["lengthscales"], ARD=True),
                                noise_variance=h["noise_variance"])
        acq = ExpectedImprovement(mk(Y[:, [0]]))
        if joint:
            acq = acq * ProbabilityOfFeasibility(mk(Y[:, [1]]))
        x = synthetic.candidates(1, Ds)
        rec = {"acquisition": name, "M": Ms, "D": Ds, "N": 1,
               "evaluate_ms": wall_ms(lambda: acq.evaluate(x)),
               "evaluate_with_gradients_ms": wall_ms(lambda: acq.evaluate_with_gradients(x))}
        t = acq._get_engine().ctx.timings()
        rec["device_ms"] = t["eval_ms"]
        rec["launches"] = int(t["launches"])
        rows.append(rec)
        acq._get_engine().close()
    return {"workload": "one candidate through Acquisition.evaluate / evaluate_with_gradients (the L-BFGS-B polish calls)",
            "timing": "wall clock per call, mean of 200, host ndarray in and out", "rows": rows}


_REAL_STDOUT = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout: everything else a library prints on fd 1 (NCCL prints its version banner
    there) is sent to stderr; emit() writes the line to the original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(obj) + "\n").encode())


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--candidates", type=int, default=N_PER_GPU)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline only: skip the configs / bars sub-records")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as td
    import gpflowopt_b200 as gpfo
    from gpflowopt_b200 import dist, synthetic
    from gpflowopt_b200.acquisition import ExpectedImprovement

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    gpfo.set_device(local_rank)
    if world > 1:
        td.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    # ---- model: the same synthetic GP on every rank (factorised redundantly; deterministic kernels) ----
    X, Y = synthetic.training_set(M, D)
    h = synthetic.hypers(D)
    model = gpfo.GPR(X, Y, gpfo.Matern52(D, variance=h["variance"], lengthscales=h["lengthscales"], ARD=True),
                     noise_variance=h["noise_variance"])
    acq = ExpectedImprovement(model)
    acq.optimize_restarts = 0
    engine = acq._get_engine()
    ctx = engine.ctx
    if args.variant:
        ctx.set_variant(args.variant)
    n_local = args.candidates
    # per-rank shard of the global candidate matrix (global row = rank * n_local + local row)
    host = torch.empty((n_local, D), dtype=torch.float64, pin_memory=True)
    host.numpy()[:] = synthetic.candidates(n_local, D, seed=4321 + rank)
    dev = host.cuda()
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")     # > 126 MB L2
    acq.evaluate(host.numpy()[:64])                       # runs setup (fmin) and uploads the program

    def step_device():
        v, i = engine.argmax_device(acq, dev.data_ptr(), n_local, D)
        vals, idxs = dist.allgather_pair(v, i + rank * n_local)
        return dist.reduce_pairs(vals, idxs)

    def step_host():
        v, i = acq.argmax(host.numpy())
        vals, idxs = dist.allgather_pair(v, i + rank * n_local)
        return dist.reduce_pairs(vals, idxs)

    for _ in range(max(args.warmup, 3)):
        flush.fill_(1.0)
        torch.cuda.synchronize()
        step_device()
    fp64_peak = ctx.measure_dmma_peak()

    # ---- timed region 1: inputs resident in HBM ----------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    dev_ms, hot_ms, launches = 0.0, 0.0, 0
    t0 = time.perf_counter()
    for s in range(args.steps):
        flush.fill_(float(s))                             # L2 flush between timed iterations
        torch.cuda.synchronize()
        best = step_device()
        t = ctx.timings()
        dev_ms += t["eval_ms"]
        hot_ms += t["hot_kernel_ms"]
        launches += t["launches"]
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.finish()

    # ---- timed region 2: end to end through the public API, host candidates ------------------------------
    for _ in range(2):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        best_e2e = step_host()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    assert best_e2e == best, "host and device candidate paths disagree: %r vs %r" % (best_e2e, best)

    variant_used = int(ctx.timings()["variant"])

    # ---- side measurements (rank-local, short): e2e from a PAGEABLE ndarray (what RandomDesign.generate() hands over), and
    # the previous default kernel (64-candidate tiles + global K(X*,X) scratch) for comparison
    pageable = np.array(host.numpy())                    # ordinary malloc'ed copy
    acq.argmax(pageable)
    t0 = time.perf_counter()
    for s in range(args.steps):
        acq.argmax(pageable)
    pageable_ms = (time.perf_counter() - t0) * 1e3
    alt_ms, alt_variant = 0.0, (12 if variant_used != 12 else 20)
    if not args.variant:
        ctx.set_variant(alt_variant)
        step_device()
        for s in range(args.steps):
            flush.fill_(float(s))
            torch.cuda.synchronize()
            step_device()
            alt_ms += ctx.timings()["eval_ms"]
        ctx.set_variant(0)

    stats = torch.tensor([dev_ms, hot_ms, e2e_ms, wall_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        td.all_reduce(stats, op=td.ReduceOp.MAX)
    dev_ms, hot_ms, e2e_ms, wall_ms = [float(v) for v in stats.cpu()]

    if rank == 0:
        n_total = n_local * world
        ms_per_step = dev_ms / args.steps
        flops_per_launch = float(n_local) * M * (M + 1)                 # algorithmic (triangular) flop, G = 1
        achieved = flops_per_launch / (hot_ms / args.steps * 1e-3) / 1e12
        out = {
            "metric": METRIC, "value": n_total / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": _workload_text(n_local), "M": M, "D": D, "candidates_per_gpu": n_local,
                       "kernel": "Matern52-ARD", "acquisition": "ExpectedImprovement",
                       "l2": "flushed by a 256 MiB device write between timed steps",
                       "parallelism": "candidate rows sharded, dp%d" % world,
                       "timing": "CUDA events on the library stream (gpfo_timings.eval_ms), max over ranks"},
            "wall_ms_per_step": wall_ms / args.steps,
            "e2e": {"value": n_total / (e2e_ms / args.steps * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": n_local * D * 8, "d2h_bytes_per_step": 16,
                    "api": "ExpectedImprovement.argmax(host ndarray) = fused MCOptimizer scoring + argmin"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak, "traffic": KERNELS.get(variant_used, {}).get("traffic"),
                         "traffic_note": "dram__bytes_read+write of ONE launch of this kernel at this workload, from the ncu "
                                         "capture named in traffic_source (not measurable outside a profiler); algorithmic "
                                         "bytes per launch = 8(D+1) N + 8 M(M+1)/2 = %d" % (8 * (D + 1) * n_local + 4 * M * (M + 1)),
                         "traffic_source": KERNELS.get(variant_used, {}).get("traffic_source"),
                         "kernel_variant": variant_used,
                         "kernel": KERNELS.get(variant_used, {}).get("kernel", "variant %d" % variant_used),
                         "flop_per_candidate": M * (M + 1),
                         "peak_source": "FP64 DMMA issue peak measured live on this GPU (gpfo_measure_dmma_peak); "
                                        "MEASURED_PEAKS.json holds no FP64 figure; nominal 37.2 TFLOP/s"},
            "best": {"value": best[0], "index": best[1]},
        }
        out["e2e_pageable"] = {"value": n_local / (pageable_ms / args.steps * 1e-3), "unit": UNIT,
                               "note": "rank 0 alone, Acquisition.argmax(pageable ndarray): the H2D copy goes through the driver's staging"}
        if alt_ms > 0.0:
            out["alt_kernel"] = {"kernel_variant": alt_variant, "value": n_local / (alt_ms / args.steps * 1e-3), "unit": UNIT,
                                 "kernel": KERNELS.get(alt_variant, {}).get("kernel"),
                                 "traffic": KERNELS.get(alt_variant, {}).get("traffic"),
                                 "note": "rank 0, same workload through the other large-batch kernel (gpfo_set_variant)"}
        if not args.no_cpu_baseline and world == 1:      # reported baseline: rank 0 at N = 1 only (the host is the same at every N)
            n_cpu = 100000
            v_cpu, dt = cpu_oracle_throughput(n_cpu)
            out["cpu_baseline"] = {"value": v_cpu, "unit": UNIT, "cores": host_threads(), "kind": "port",
                                   "sample": "%d candidates (%.1f s) in 4096-row chunks, Cholesky hoisted, "
                                             "NumPy/SciPy + OpenBLAS on all host cores" % (n_cpu, dt)}
    engine.close()
    del dev, flush
    torch.cuda.empty_cache()
    if not args.no_extras:
        t_extra = time.perf_counter()
        cfg3 = sub_config3_constrained(torch, td, gpfo, world, rank, barrier)          # collective: every rank takes part
        if rank == 0:
            out["configs"] = {"configs[2]": cfg3}
            if world == 1:                                # single-GPU records and bars: once, in the N = 1 run
                out["configs"]["configs[0]"] = sub_config1_branin(gpfo)
                out["configs"]["configs[3]"] = sub_config4_hvpoi(gpfo)
                out["configs"]["configs[4]"] = sub_config5_sweep(gpfo, fp64_peak)
                out["polish_latency"] = sub_polish_latency(gpfo)
                out["bars"] = gpu_library_bars(torch, fp64_peak)
                if not args.no_cpu_baseline:
                    out["bars"].update(cpu_side_bars())
            out["extras_wall_s"] = time.perf_counter() - t_extra
        barrier()
    if rank == 0:
        emit(out)
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
