python tools/grad_prof.py 75776 2048 > gpurun_out/r2_plain_grad.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:contract -c 2 -o gpurun_out/r2_prof_grad -f python tools/grad_prof.py 75776 2048 > gpurun_out/r2_ncu_grad.log 2>&1
cat gpurun_out/r2_plain_grad.log
python tools/prof_run.py 0 1000000 2048 2 > gpurun_out/r2_plain_1e6.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:contract5 -s 1 -c 1 --csv --log-file gpurun_out/r2_v20_dram_1e6.csv python tools/prof_run.py 0 1000000 2048 2 > gpurun_out/r2_ncu_1e6.log 2>&1
cat gpurun_out/r2_plain_1e6.log; tail -5 gpurun_out/r2_v20_dram_1e6.csv
python tools/hvpoi_prof.py > gpurun_out/r2_plain_hv.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:hvpoi2 -c 1 -o gpurun_out/r2_prof_hvpoi2 -f python tools/hvpoi_prof.py > gpurun_out/r2_ncu_hv.log 2>&1
tail -2 gpurun_out/r2_plain_hv.log
