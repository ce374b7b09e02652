"""BASELINE configs[2]: constrained BO scoring, EI x PoF (2 GPs, M = 4096, D = 16) over 1e7 candidates sharded across the
GPUs of one box.   torchrun --nproc-per-node G tools/config3_dist.py [N]
Every rank factorises both GPs redundantly, scores its contiguous row block (host candidates: MCOptimizer as in the
reference; then the same job with device-generated candidates) and the ranks all-gather one (value, index) record."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as td

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpflowopt_b200 as gpfo                                   # noqa: E402
from gpflowopt_b200 import dist, synthetic                      # noqa: E402
from gpflowopt_b200.acquisition import ExpectedImprovement, ProbabilityOfFeasibility   # noqa: E402
from gpflowopt_b200.bo import _InverseAcquisition               # noqa: E402
from gpflowopt_b200.domain import UnitCube                      # noqa: E402
from gpflowopt_b200.optim import MCOptimizer                    # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
gpfo.set_device(local)
td.init_process_group("nccl", device_id=torch.device("cuda", local))
M, D = 4096, 16
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10 ** 7
X, Y = synthetic.training_set(M, D, constraint=True)
h = synthetic.hypers(D)
mk = lambda c: gpfo.GPR(X, Y[:, [c]], gpfo.Matern52(D, variance=h["variance"], lengthscales=h["lengthscales"], ARD=True),
                        noise_variance=h["noise_variance"])
t0 = time.perf_counter()
joint = ExpectedImprovement(mk(0)) * ProbabilityOfFeasibility(mk(1))
joint.evaluate(np.zeros((1, D)) + 0.5)                           # setup: factorisations, fmin, program upload
t_setup = time.perf_counter() - t0
dom = UnitCube(D)
for label, opt in (("host candidates (reference MCOptimizer semantics)", MCOptimizer(dom, N)),
                   ("device-generated candidates", MCOptimizer(dom, N, device_rng=True))):
    np.random.seed(4321)
    td.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = opt.optimize(_InverseAcquisition(joint, False))
    torch.cuda.synchronize()
    td.barrier()
    dt = time.perf_counter() - t0
    dev_ms = joint._get_engine().ctx.timings()["eval_ms"]
    tmax = torch.tensor([dt, dev_ms], dtype=torch.float64, device="cuda")
    td.all_reduce(tmax, op=td.ReduceOp.MAX)
    check = -joint.evaluate(res.x)[0, 0]                          # the winning row, scored alone, reproduces the value
    xs = [None] * world
    td.all_gather_object(xs, (res.x.tobytes(), float(np.ravel(res.fun)[0])))
    assert all(v == xs[0] for v in xs), "ranks disagree on the selected point"
    assert abs(check - float(np.ravel(res.fun)[0])) <= 1e-12 * abs(check), (check, res.fun)
    if rank == 0:
        print("config3 %d GPUs, N=%d, %s: wall %.2f s (%.3e cand/s incl. candidate generation/transfer), device eval %.1f ms max over ranks "
              "(%.3e cand/s, %.2f TF/s per GPU); best acq %.6g; setup (2 factorisations + fmin) %.2f s" % (
                  world, N, label, tmax[0].item(), N / tmax[0].item(), tmax[1].item(), N / (tmax[1].item() * 1e-3),
                  2.0 * (N / world) * M * (M + 1.0) / (tmax[1].item() * 1e-3) / 1e12, -float(np.ravel(res.fun)[0]), t_setup), flush=True)
td.destroy_process_group()
