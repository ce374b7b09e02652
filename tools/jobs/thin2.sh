timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "latency_path or row_split or mes_leaf or leaf_gradients or product_sum or slot_offsets or api_errors or argmax_first" > gpurun_out/r2_thin_tests.log 2>&1; echo "thin tests rc $?"; tail -4 gpurun_out/r2_thin_tests.log
timeout 600 python tools/latency.py > gpurun_out/r2_latency_thin2.log 2>&1; grep -E "N=    1|N=   32|N= 1000|staged|same" gpurun_out/r2_latency_thin2.log
GPFO_THIN_MAX_N=0 timeout 600 python tools/latency.py 2>&1 | grep -E "N=    1:|N=   32:" > gpurun_out/r2_latency_thin_oldkernels.log; cat gpurun_out/r2_latency_thin_oldkernels.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_suite7.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/r2_gpu_suite7.log
