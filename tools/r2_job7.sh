#!/bin/bash
# 1 GPU: MPC launch geometry A/B (one agent-warp per CTA vs packed CTAs) on the BASELINE MPC shapes + config 5
O=gpurun_out/r2; mkdir -p $O
for g in "" 2 8; do
  echo "== SCB_MPC_GPB=${g:-default(1)}"
  SCB_MPC_GPB=$g timeout 300 python tools/mpc_variants.py cfg3 du5 kb5 q5 2>&1 | tail -4
  SCB_MPC_GPB=$g timeout 300 python bench.py --workload cfg5 --steps 3 --warmup 1 --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg5 ms', d['ms_per_step'], 'value', d['value'])"
done 2>&1 | tee $O/mpc_gpb_ab.txt
