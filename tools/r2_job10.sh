#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
timeout 300 python tools/time_backup.py 65536 10 2>&1 | tee $O/backup_timing2.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:backup -c 60 --csv --log-file $O/launches_backup.csv python tools/time_backup.py 65536 2 > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2/launches_backup.csv")) if len(r) > 5 and r[0].isdigit()]
for r in rows[:40]: print(r[4][:40], r[-1])
PY
