timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_suite3.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/r2_gpu_suite3.log
timeout 900 python tools/grad_bench.py > gpurun_out/r2_grad_bench.log 2>&1; cat gpurun_out/r2_grad_bench.log
GPFO_GRAD_FWD_V12=1 timeout 900 python tools/grad_bench.py > gpurun_out/r2_grad_bench_fwd12.log 2>&1; grep stored gpurun_out/r2_grad_bench_fwd12.log
