#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
(time timeout 1700 python -m pytest tests -m gpu -x -q) > $O/pytest_gpu.log 2>&1; tail -8 $O/pytest_gpu.log
timeout 1200 bash tools/prof_r2.sh
timeout 1500 bash tools/sanitize.sh
