#!/bin/bash
# refresh the MPC evidence with the final kernels: ncu --set full at the cfg3 shape + counted flops of every MPC workload
O=gpurun_out/r2; mkdir -p $O
timeout 900 bash tools/prof_mpc_ncu.sh 4096 > $O/ncu_mpc.log 2>&1
cp gpurun_out/p/mpc_summary.txt $O/ncu_mpc_summary_v2.txt; cp gpurun_out/p/mpc_stalls.txt $O/ncu_mpc_stalls_v2.txt; cp gpurun_out/p/mpc_hot_lines.txt $O/ncu_mpc_hot_lines_v2.txt
head -36 $O/ncu_mpc_summary_v2.txt
timeout 600 bash tools/count_flops.sh
python -c "
import json; d = json.load(open('gpurun_out/r2/r2_flops.json')); print({k: v for k, v in d.items() if not k.startswith('_')})"
