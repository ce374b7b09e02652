import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflowopt_b200 import _lib
ctx = _lib.Context(0)
for mode, name in ((0, "DMMA only"), (1, "DFMA only"), (2, "DMMA + DFMA concurrently (2+2 warps per scheduler)")):
    a, b = ctx.measure_fp64_pipes(mode)
    print("%-55s DMMA %.2f TF/s  DFMA %.2f TF/s" % (name, a, b))
