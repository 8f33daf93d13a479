#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
(time timeout 900 python -m pytest tests/test_gpu_backup.py -x -q -m gpu) > $O/pytest_gpu_backup.log 2>&1; tail -6 $O/pytest_gpu_backup.log
timeout 300 python tools/time_backup.py 65536 10 2>&1 | tee $O/backup_timing3.txt
timeout 300 python tools/time_backup.py 2048 10 2>&1 | tee -a $O/backup_timing3.txt
