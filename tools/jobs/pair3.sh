export GPFO_BUILD_EXPERIMENTS=1
python tools/variant_sweep.py 21,37,29 512 > gpurun_out/r2_nw8_a.log 2>&1; tail -4 gpurun_out/r2_nw8_a.log
python tools/variant_sweep.py 20,36,29 1024,2048 > gpurun_out/r2_nw8_b.log 2>&1; tail -7 gpurun_out/r2_nw8_b.log
