python tools/grad_prof.py 75776 2048 > gpurun_out/r2_plain_grad2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:contract5 -c 2 -o gpurun_out/r2_prof_grad2 -f python tools/grad_prof.py 75776 2048 > gpurun_out/r2_ncu_grad2.log 2>&1
cat gpurun_out/r2_plain_grad2.log
