#!/bin/bash
# FP64 flop count of every MPC workload of the bench -> gpurun_out/r2/flops.csv (+ r2_flops.json via tools/count_flops.py)
O=gpurun_out/r2; mkdir -p $O
ncu --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum \
    --clock-control none -k regex:mpc_kernel --csv --log-file $O/flops.csv python tools/prof_flops.py > $O/flops.log 2>&1
python tools/count_flops.py $O/flops.csv $O/r2_flops.json >> $O/flops.log 2>&1
tail -5 $O/flops.log
