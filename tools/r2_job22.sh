#!/bin/bash
# MPC: layout in the constant bank + occupancy-sized one-warp CTAs -- A/B against the 8-slot cap
O=gpurun_out/r2; mkdir -p $O
for v in "" 1; do
  echo "== SCB_MPC_SLOTS8=${v:-unset}"
  SCB_MPC_SLOTS8=$v timeout 300 python tools/mpc_variants.py cfg3 du5 kb5 q5 2>&1 | tail -4
done 2>&1 | tee $O/mpc_layout_ab.txt
timeout 300 python bench.py --workload cfg5 --steps 3 --warmup 1 --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg5 1 GPU ms', d['ms_per_step'], 'value', d['value'])" | tee -a $O/mpc_layout_ab.txt
(time timeout 900 python -m pytest tests/test_gpu_mpc.py tests/test_gpu_track.py -x -q -m gpu) 2>&1 | tail -6
