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
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel for one launch of this workload (N = 1e6),
# taken from profiles/r01_final_contract2_v12.metrics.csv (27.7 GB read + 16.3 GB written: the per-CTA K(X*,X) scratch of
# variant 12 is written back to HBM and partly re-read from it -- an earlier capture of the same kernel saw 44.7 GB read, the
# split between L2 hits and HBM varies; the strictly on-chip variant 4 moves 36.5 MB per launch)
NCU_DRAM_BYTES_PER_LAUNCH = 27688939000 + 16286433000      # profiles/r01_final_contract2_v12.metrics.csv (read + write)
UNIT = "candidates/s"


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

    # ---- side measurement: the strictly-on-chip variant (K(X*,X) never leaves the SM; no global scratch) ----------
    strict_ms = 0.0
    if not args.variant:
        ctx.set_variant(4)
        step_device()
        for s in range(args.steps):
            flush.fill_(float(s))
            torch.cuda.synchronize()
            step_device()
            strict_ms += ctx.timings()["eval_ms"]
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
                         "frac": achieved / fp64_peak, "traffic": NCU_DRAM_BYTES_PER_LAUNCH,
                         "traffic_note": "dram__bytes_read+write of one contract2_kernel launch at this workload from "
                                         "profiles/ (ncu --set full); bytes, not flop: the kernel is tensor bound, HBM is not",
                         "kernel": "contract2_kernel<16,2,8,4> (FP64 DMMA.8x8x4 contraction + fused EI epilogue + block "
                                   "argmax; 64-candidate tiles, 256-row chunks, K(X*,X) generated once per tile)",
                         "flop_per_candidate": M * (M + 1),
                         "peak_source": "FP64 DMMA issue peak measured live on this GPU (gpfo_measure_dmma_peak); "
                                        "MEASURED_PEAKS.json holds no FP64 figure; nominal 37.2 TFLOP/s"},
            "best": {"value": best[0], "index": best[1]},
        }
        if strict_ms > 0.0:
            out["strict_onchip"] = {"value": n_local / (strict_ms / args.steps * 1e-3), "unit": UNIT,
                                    "note": "rank 0, variant 4: covariance tiles strictly on chip (no global scratch), "
                                            "the literal north_star dataflow"}
        if not args.no_cpu_baseline:
            n_cpu = 100000
            v_cpu, dt = cpu_oracle_throughput(n_cpu)
            out["cpu_baseline"] = {"value": v_cpu, "unit": UNIT, "cores": host_threads(), "kind": "port",
                                   "sample": "%d candidates (%.1f s) in 4096-row chunks, Cholesky hoisted, "
                                             "NumPy/SciPy + OpenBLAS on all host cores" % (n_cpu, dt)}
        emit(out)
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
