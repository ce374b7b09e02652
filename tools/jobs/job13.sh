GPFO_NO_GRAPHS=1 python tools/n1_prof.py && GPFO_NO_GRAPHS=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_n1_launches.csv python tools/n1_prof.py > gpurun_out/r2_n1_ncu.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2_n1_launches.csv")) if len(r) > 10 and r[0].isdigit()]
for r in rows[-14:]:
    print(r[4][:90], r[7], r[8], r[-1])
PY
