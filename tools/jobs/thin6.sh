timeout 600 python tools/latency.py > gpurun_out/r2_latency_thin4.log 2>&1; grep -E "N=  |staged|same" gpurun_out/r2_latency_thin4.log
GPFO_THIN_MAX_N=0 timeout 600 python tools/latency.py 2>&1 | grep -E "N=    1:|N=    4:|N=    8:" > gpurun_out/r2_latency_thin_oldkernels.log; cat gpurun_out/r2_latency_thin_oldkernels.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "latency_path or row_split or mes_leaf or leaf_gradients or product_sum or slot_offsets" > gpurun_out/r2_thin_tests.log 2>&1; echo "thin tests rc $?"; tail -2 gpurun_out/r2_thin_tests.log
