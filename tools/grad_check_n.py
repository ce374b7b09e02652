"""Stored-A gradient pass at M = 4096: spread of repeated calls (device time), with the clocks sampled alongside.
   python tools/grad_check_n.py [M] [N]"""
import os, subprocess, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflowopt_b200 import _lib, synthetic
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
ctx = _lib.Context(0)
peak = ctx.measure_dmma_peak()
X, Y = synthetic.training_set(M, 8)
ctx.gp_set(0, X, Y, _lib.KERN_MATERN52, **synthetic.hypers(8))
ctx.program_set([[_lib.OP_EI, 0, 0, 0]], [-1.5])
Xc = synthetic.candidates(N, 8)
ctx.eval(Xc, want_grads=True)
ts = []
for i in range(12):
    ctx.eval(Xc, want_grads=True)
    t = ctx.timings()
    clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader"],
                         capture_output=True, text=True).stdout.strip()
    ts.append(t["eval_ms"])
    print("call %2d: %.1f ms  %.1f %% of peak  launches %d   [%s]" % (i, t["eval_ms"], 100 * N * 2.0 * M * (M + 1.0) / (t["eval_ms"] * 1e-3) / 1e12 / peak,
                                                                    t["launches"], clk), flush=True)
print("min %.1f median %.1f max %.1f ms" % (min(ts), float(np.median(ts)), max(ts)))
