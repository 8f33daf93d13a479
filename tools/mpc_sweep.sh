#!/bin/bash
# one gpurun call: MPC kernel variants on the BASELINE shapes + per-phase cycle profile
#   VARIANTS="'' _t384" CASES="cfg3 du5" bash tools/mpc_sweep.sh
O=gpurun_out/mpc_sweep; mkdir -p $O; : > $O/variants.txt
for v in ${VARIANTS:-""}; do
  [ "$v" = "base" ] && v=""
  for sch in ${SCHED:-1}; do
  SCB_MPC_SCHEDULE=$sch SCB_LIB=$PWD/safe_control_b200/libscb$v.so timeout 300 python tools/mpc_variants.py ${CASES:-cfg3 du5 kb5 q5 si} >> $O/variants.txt 2>&1
  done
done
[ -f safe_control_b200/libscb_prof.so ] && SCB_LIB=$PWD/safe_control_b200/libscb_prof.so timeout 300 python tools/prof_mpc_phases.py > $O/phases.txt 2>&1
cat $O/variants.txt $O/phases.txt
