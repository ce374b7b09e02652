export GPFO_BUILD_EXPERIMENTS=1
for v in 20 22 23 24 25 26; do python tools/prof_run.py $v 151552 2048 2 2>&1 | tail -1; done > gpurun_out/r2_v5_ablate.log; cat gpurun_out/r2_v5_ablate.log
python tools/phase_timing5.py 27 2048 > gpurun_out/r2_phase5_v20_M2048.log 2>&1; cat gpurun_out/r2_phase5_v20_M2048.log
python tools/phase_timing5.py 28 512 > gpurun_out/r2_phase5_v21_M512.log 2>&1; cat gpurun_out/r2_phase5_v21_M512.log
python tools/phase_timing5.py 27 1024 > gpurun_out/r2_phase5_v20_M1024.log 2>&1; cat gpurun_out/r2_phase5_v20_M1024.log
