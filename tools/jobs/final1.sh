python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc $?"; tail -3 gpurun_out/r2_final_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_gpu_suite.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/r2_final_gpu_suite.log
timeout 900 python bench.py > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err; echo "bench rc $?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_final_bench_n1.json"))
print("value %.4g e2e %.4g frac %.3f variant %s traffic %s pageable %.4g alt %.4g cpu %.4g" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_variant"], d["roofline"]["traffic"], d["e2e_pageable"]["value"], d["alt_kernel"]["value"], d["cpu_baseline"]["value"]))
c = d["configs"]
print("cfg2", {k: round(c["configs[2]"][k]["wall_s"], 2) for k in ("host_candidates", "device_rng")})
print("cfg3", {k: c["configs[3]"][k] for k in ("eval_ms", "gp_contractions_ms", "cell_kernel_ms", "pareto_front")})
for r in c["configs[4]"]["rows"]:
    print("cfg4 M=%d values %.3f (v%d) grads %.3f" % (r["M"], r["values"]["frac_of_dmma_peak"], r["values"]["kernel_variant"], r["gradients"]["frac_of_dmma_peak"]))
print({k: (round(v["value"], 3)) for k, v in d["bars"].items()}, "extras %.1f s" % d["extras_wall_s"])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_bench_reference_arm.json 2>/dev/null; cut -c1-300 gpurun_out/r2_final_bench_reference_arm.json
