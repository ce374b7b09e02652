export GPFO_BUILD_EXPERIMENTS=1
python tools/variant_sweep.py 21,38,39 512,1024 > gpurun_out/r2_abl21.log 2>&1; tail -7 gpurun_out/r2_abl21.log
