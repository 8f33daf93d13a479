#!/bin/bash
# compute-sanitizer memcheck + racecheck over a small invocation of every kernel family;
# logs -> gpurun_out/r2/sanitizer_{memcheck,racecheck}.log (summaries are committed under profiles/)
O=gpurun_out/r2; mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/sanitizer_memcheck.log
SAN_N=2400 timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 python tools/sanitize_smoke.py > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/sanitizer_racecheck.log
tail -4 $O/sanitizer_memcheck.log; tail -6 $O/sanitizer_racecheck.log
