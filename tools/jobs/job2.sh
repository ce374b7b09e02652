timeout 600 python -m pytest tests -m gpu -x -q -k "hvpoi or config4 or vlmop or hv" > gpurun_out/r2_hv_tests.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r2_hv_tests.log
python - <<'PY' > gpurun_out/r2_hv_bench.log 2>&1
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import bench, gpflowopt_b200 as gpfo
for env in ("1", ""):
    if env: os.environ["GPFO_HVPOI_V1"] = env
    else: os.environ.pop("GPFO_HVPOI_V1", None)
    r = bench.sub_config4_hvpoi(gpfo)
    print("GPFO_HVPOI_V1=%r: eval %.2f ms, contractions %.2f, cells %.2f ms, P=%s, best %r" % (env, r["eval_ms"], r["gp_contractions_ms"], r["cell_kernel_ms"], r["pareto_front"], r["best"]))
PY
cat gpurun_out/r2_hv_bench.log
GPFO_BUILD=1 true
for v in 20 22 23 24 25 26; do python tools/prof_run.py $v 151552 2048 2 2>&1 | tail -1; done > gpurun_out/r2_v5_ablate.log; cat gpurun_out/r2_v5_ablate.log
