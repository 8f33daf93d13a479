#!/bin/bash
# 1 GPU: ncu --set full of the Backup-CBF kernels (split rollout + QP, fused), 8 lanes per agent, 65 536 agents
O=gpurun_out/r2; P=gpurun_out/p; mkdir -p $O $P
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'backup' -s 0 -c 3 -f -o $P/bk python tools/prof_backup.py 65536 > $P/bk.log 2>&1
python tools/ncu_summary.py $P/bk.ncu-rep --title "backup-cbf (round 2): python tools/prof_backup.py 65536 -- backup_rollout_kernel<8>, backup_qp_kernel<8,16>, backupcbf_kernel<8,16> (fused)" > $O/ncu_backup_summary.txt 2>> $P/bk.log
ncu -i $P/bk.ncu-rep --page source --csv 2>/dev/null | python tools/ncu_hot_lines.py > $O/ncu_backup_hot_lines.txt 2>> $P/bk.log
rm -f $P/bk.ncu-rep
tail -3 $P/bk.log; cat $O/ncu_backup_summary.txt | head -150
