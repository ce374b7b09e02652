python tools/thin_debug.py 2048 8 2>&1 | tail -18 > gpurun_out/r2_thin_debug.log; cat gpurun_out/r2_thin_debug.log
python tools/thin_debug.py 4096 16 2>&1 | tail -18 > gpurun_out/r2_thin_debug4096.log; cat gpurun_out/r2_thin_debug4096.log
