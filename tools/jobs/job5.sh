export GPFO_BUILD_EXPERIMENTS=1
python tools/variant_sweep.py 21,31,33,35 512,1024 > gpurun_out/r2_sched_a.log 2>&1; tail -12 gpurun_out/r2_sched_a.log
python tools/variant_sweep.py 20,30,32,34 1024,2048 > gpurun_out/r2_sched_b.log 2>&1; tail -12 gpurun_out/r2_sched_b.log
