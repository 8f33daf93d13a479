#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -k regex:shield -c 12 --csv --log-file $O/launches_shield.csv python tools/prof_shield.py 65536 > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2/launches_shield.csv")) if len(r) > 6 and r[0].isdigit()]
for r in rows: print(r[4][:44], r[-3], r[-1])
PY
