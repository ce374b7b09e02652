timeout 600 python tools/latency.py > gpurun_out/r2_latency_thin3.log 2>&1; grep -E "N=  |staged|same" gpurun_out/r2_latency_thin3.log
GPFO_THIN_MAX_N=0 timeout 600 python tools/latency.py 2>&1 | grep -E "N=    4:|N=    8:" > gpurun_out/r2_latency_thin_oldkernels.log; cat gpurun_out/r2_latency_thin_oldkernels.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_suite7.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/r2_gpu_suite7.log
