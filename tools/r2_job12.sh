#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
(time timeout 1200 python -m pytest tests/test_gpu_shield.py tests/test_gpu_backup.py -x -q -m gpu) > $O/pytest_gpu_shield.log 2>&1
tail -30 $O/pytest_gpu_shield.log
timeout 300 python tools/time_shield.py 65536 2>&1 | tee $O/shield_timing.txt
