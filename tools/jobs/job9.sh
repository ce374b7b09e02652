NG=${NG:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/r2_dist_check_n$NG.log 2>&1; echo "dist_check rc $?"; grep -E "ok:|Error|assert" gpurun_out/r2_dist_check_n$NG.log | head -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 5 --warmup 3 > gpurun_out/r2_bench_n$NG.json 2> gpurun_out/r2_bench_n$NG.err; echo "bench rc $?"
NG=$NG python - <<'PY'
import json, os
ng = os.environ["NG"]
d = json.load(open("gpurun_out/r2_bench_n%s.json" % ng))
print("value %.4g  e2e %.4g  frac %.3f ms/step %.2f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["ms_per_step"]))
c = d["configs"]["configs[2]"]
for k in ("host_candidates", "device_rng"):
    print(k, {kk: c[k][kk] for kk in ("wall_s", "device_ms", "value_wall", "tflops_per_gpu", "scoring_calls_per_rank", "best_acq")})
print("extras", d.get("extras_wall_s"))
PY
