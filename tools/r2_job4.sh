#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
(time timeout 1700 python -m pytest tests -m gpu -x -q) > $O/pytest_gpu.log 2>&1; tail -8 $O/pytest_gpu.log
for v in "" _mb5 _mb4; do echo "== libscb$v" >> $O/sweep_minblocks.txt; SCB_LIB=$PWD/safe_control_b200/libscb$v.so SWEEP_ONLY=big timeout 300 python tools/sweep_tma.py >> $O/sweep_minblocks.txt 2>&1; done; cat $O/sweep_minblocks.txt
SAN_N=2400 timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 python tools/sanitize_smoke.py > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/sanitizer_racecheck.log; tail -4 $O/sanitizer_racecheck.log
(time timeout 900 python bench.py --steps 20 --warmup 5) > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 300 $O/bench_n1.err
