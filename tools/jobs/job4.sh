export GPFO_BUILD_EXPERIMENTS=1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "leaf_acquisitions or multi_chunk or resident_panel or config5 or headline or predict_parity or random or device" > gpurun_out/r2_v5b_tests.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/r2_v5b_tests.log
timeout 600 python tools/variant_sweep.py 12,20,21 512,1024,2048,4096,8192 > gpurun_out/r2_v5_sweep4.log 2>&1; cat gpurun_out/r2_v5_sweep4.log
python tools/phase_timing5.py 27 2048 > gpurun_out/r2_phase5b_v20_M2048.log 2>&1; head -6 gpurun_out/r2_phase5b_v20_M2048.log
python tools/phase_timing5.py 28 512 > gpurun_out/r2_phase5b_v21_M512.log 2>&1; head -6 gpurun_out/r2_phase5b_v21_M512.log
