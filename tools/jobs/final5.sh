bash tools/jobs/final1.sh
timeout 600 python tools/latency.py > gpurun_out/r2_latency_final.log 2>&1; grep -E "N=    1:|N=    4:|N= 1000:|EI x PoF|staged|same|MC\(1e6\)" gpurun_out/r2_latency_final.log
