timeout 600 python tools/latency.py > gpurun_out/r2_latency_thin.log 2>&1; grep -E "N=    1|N=   32|staged|same" gpurun_out/r2_latency_thin.log
GPFO_THIN_COPIES=1 timeout 600 python tools/latency.py 2>&1 | grep -E "N=    1:" > gpurun_out/r2_latency_thin_copies.log; cat gpurun_out/r2_latency_thin_copies.log
GPFO_NO_THIN=1 timeout 600 python tools/latency.py 2>&1 | grep -E "N=    1:" > gpurun_out/r2_latency_nothin.log; cat gpurun_out/r2_latency_nothin.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_suite7.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/r2_gpu_suite7.log
