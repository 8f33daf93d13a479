#!/bin/bash
O=gpurun_out/r2; P=gpurun_out/p; mkdir -p $O $P
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'shield_step' -s 1 -c 1 -f -o $P/sh python tools/prof_shield.py 65536 > $P/sh.log 2>&1
python tools/ncu_summary.py $P/sh.ncu-rep --title "gatekeeper (round 2): python tools/prof_shield.py 65536 -- second step of shield_step_kernel<8> (every agent re-plans)" > $O/ncu_shield_summary.txt 2>> $P/sh.log
ncu -i $P/sh.ncu-rep --page source --csv 2>/dev/null | python tools/ncu_hot_lines.py > $O/ncu_shield_hot_lines.txt 2>> $P/sh.log
rm -f $P/sh.ncu-rep
tail -3 $P/sh.log; cat $O/ncu_shield_summary.txt | head -60; head -30 $O/ncu_shield_hot_lines.txt
