timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "grad or stored or config5 or slab" > gpurun_out/r2_bwd5_tests.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/r2_bwd5_tests.log
timeout 900 python tools/grad_bench.py 2>&1 | grep "stored\|peak" > gpurun_out/r2_grad_bench_bwd5.log; cat gpurun_out/r2_grad_bench_bwd5.log
GPFO_GRAD_BWD_V2=1 timeout 900 python tools/grad_bench.py 2>&1 | grep "stored" 
