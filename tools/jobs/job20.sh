timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_suite6.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/r2_gpu_suite6.log
timeout 600 python tools/variant_sweep.py 20,21 512,1024,2048,4096,8192 > gpurun_out/r2_v5_sweep5.log 2>&1; cat gpurun_out/r2_v5_sweep5.log
timeout 900 python tools/grad_bench.py 2>&1 | grep "stored"
