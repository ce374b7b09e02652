timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_suite4.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r2_gpu_suite4.log
timeout 600 python tools/latency.py > gpurun_out/r2_latency_graphs.log 2>&1; cat gpurun_out/r2_latency_graphs.log
GPFO_NO_GRAPHS=1 timeout 600 python tools/latency.py 2>&1 | grep "N=    1:\|N=   32:\|N= 1000" > gpurun_out/r2_latency_nographs.log; cat gpurun_out/r2_latency_nographs.log
python tools/determinism_check.py 3 2>&1 | tail -3
