timeout 900 python bench.py > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err; echo "bench rc $?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_final_bench_n1.json"))
print("value %.4g e2e %.4g frac %.3f variant %s traffic %s pageable %.4g alt %.4g cpu %.4g" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_variant"], d["roofline"]["traffic"], d["e2e_pageable"]["value"], d["alt_kernel"]["value"], d["cpu_baseline"]["value"]))
c = d["configs"]
print("cfg3", {k: c["configs[3]"][k] for k in ("eval_ms", "gp_contractions_ms", "cell_kernel_ms", "pareto_front")})
for r in c["configs[4]"]["rows"]:
    print("cfg4 M=%d values %.3f (v%d) grads %.3f" % (r["M"], r["values"]["frac_of_dmma_peak"], r["values"]["kernel_variant"], r["gradients"]["frac_of_dmma_peak"]))
print("extras %.1f s" % d["extras_wall_s"], d["clocks"])
PY
python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r2_final_plain_short.json 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r2_final_ncu_launches.log 2>&1; echo "launch list rc $?"
python tools/prof_run.py 20 75776 2048 2 > gpurun_out/r2_final_plain20.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:contract5_kernel -s 1 -c 1 -o gpurun_out/r2_final_prof_v20 -f python tools/prof_run.py 20 75776 2048 2 > gpurun_out/r2_final_ncu20.log 2>&1; echo "ncu full rc $?"; cat gpurun_out/r2_final_plain20.log
