export GPFO_BUILD_EXPERIMENTS=1
python tools/variant_sweep.py 21,37 512,1024 > gpurun_out/r2_pair_a.log 2>&1; tail -5 gpurun_out/r2_pair_a.log
python tools/variant_sweep.py 20,36 1024,2048,4096 > gpurun_out/r2_pair_b.log 2>&1; tail -7 gpurun_out/r2_pair_b.log
