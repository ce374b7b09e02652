timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_suite5.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/r2_gpu_suite5.log
timeout 900 python tools/grad_bench.py > gpurun_out/r2_grad_bench_final.log 2>&1; grep stored gpurun_out/r2_grad_bench_final.log
python tools/grad_prof.py 75776 2048 > gpurun_out/r2_plain_grad3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:contract5 -c 2 -o gpurun_out/r2_prof_grad3 -f python tools/grad_prof.py 75776 2048 > gpurun_out/r2_ncu_grad3.log 2>&1
cat gpurun_out/r2_plain_grad3.log
python tools/grad_mid2.py > gpurun_out/r2_grad_mid3.log 2>&1; grep "M=2048" gpurun_out/r2_grad_mid3.log
