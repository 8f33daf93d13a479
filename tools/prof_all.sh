#!/bin/bash
# Profile batch for one gpurun call: launch lists + `ncu --set full` captures of every kernel family, summarised on the
# GPU box (tools/ncu_summary.py) so that only text comes back (the binary reports exceed gpurun_out's 64 MiB cap).
#   gpurun --timeout 1500 -- 'bash tools/prof_all.sh'
set -x
O=gpurun_out/p; mkdir -p $O; rm -f $O/*.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg2.csv python bench.py --steps 20 --warmup 3 --no-cpu > $O/launches_cfg2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_loop2.csv python bench.py --workload loop2 --steps 20 --warmup 3 --no-cpu > $O/launches_loop2.log 2>&1
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o $O/$name "$@" > $O/$name.log 2>&1
  python tools/ncu_summary.py $O/$name.ncu-rep --title "$name: $*" > $O/${name}_summary.txt 2>> $O/$name.log
  ncu -i $O/$name.ncu-rep --page source --csv 2>/dev/null | python tools/ncu_hot_lines.py > $O/${name}_hot_lines.txt 2>> $O/$name.log
  rm -f $O/$name.ncu-rep
}
cap qp 'cbfqp_kernel|odcbf_kernel' 0 9 python tools/prof_qp.py
cap mpc mpc_kernel 2 1 python tools/prof_mpc.py 1184
cap loop_fused track_ 1 2 python tools/prof_loop.py 1024 16 50
SCB_TRACK_FUSED=0 cap loop_steps 'track_|cbfqp' 9 3 python tools/prof_loop.py 1024 16 10
ls -la $O
