python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/r2_final_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_gpu_suite.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/r2_final_gpu_suite.log
bash tools/jobs/final2.sh
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_bench_reference_arm.json 2>/dev/null; cut -c1-200 gpurun_out/r2_final_bench_reference_arm.json
timeout 600 python tools/latency.py > gpurun_out/r2_latency_final.log 2>&1; grep -E "N=    1:|EI x PoF|staged|same" gpurun_out/r2_latency_final.log
