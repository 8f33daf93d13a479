#!/bin/bash
# ncu --set full captures of the round-2 kernels -> text summaries (the binary reports exceed gpurun's return cap)
O=gpurun_out/r2; P=gpurun_out/p; mkdir -p $O $P
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'cbfqp_kernel|cbfqp_tma_kernel|odcbf_tma_kernel|odcbf_kernel' -s 0 -c 9 -f -o $P/qp python tools/prof_qp.py > $P/qp.log 2>&1
python tools/ncu_summary.py $P/qp.ncu-rep --title "qp (round 2): python tools/prof_qp.py -- 3 x cbfqp_kernel<1,32,1,1> at N=1024, 3 x cbfqp_tma_kernel<1,4,5> at N=1Mi, 3 x odcbf_tma_kernel<3,1,8,4> at config 4" > $O/ncu_qp_summary.txt 2>> $P/qp.log
ncu -i $P/qp.ncu-rep --page source --csv 2>/dev/null | python tools/ncu_hot_lines.py > $O/ncu_qp_hot_lines.txt 2>> $P/qp.log
rm -f $P/qp.ncu-rep
timeout 900 bash tools/prof_mpc_ncu.sh 4096 > $O/ncu_mpc.log 2>&1
cp $P/mpc_summary.txt $O/ncu_mpc_summary.txt; cp $P/mpc_stalls.txt $O/ncu_mpc_stalls.txt; cp $P/mpc_hot_lines.txt $O/ncu_mpc_hot_lines.txt
tail -3 $P/qp.log; grep -E "^##|time_duration|dram__bytes|issue_active|warps_active|registers|local_loads|sectors_pipe|requests_pipe|fp64" $O/ncu_qp_summary.txt | head -80
