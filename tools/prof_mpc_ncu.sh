#!/bin/bash
# ncu --set full of the MPC kernel (cfg3 shape) + every warp-stall metric of the raw page
O=gpurun_out/p; mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:mpc_kernel -s 2 -c 1 -f -o $O/mpc python tools/prof_mpc.py ${1:-4096} > $O/mpc.log 2>&1
python tools/ncu_summary.py $O/mpc.ncu-rep --title "mpc: python tools/prof_mpc.py ${1:-4096}" > $O/mpc_summary.txt 2>> $O/mpc.log
ncu -i $O/mpc.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin)); h, u, d = rows[0], rows[1], rows[2]
for i, n in enumerate(h):
    if 'stall' in n or 'icc' in n or 'inst_issued' in n or 'no_instruction' in n or 'idc' in n or 'immc' in n:
        print(f'{n:100s} {d[i]:>14s} {u[i]}')
" > $O/mpc_stalls.txt
ncu -i $O/mpc.ncu-rep --page source --csv 2>/dev/null | python tools/ncu_hot_lines.py > $O/mpc_hot_lines.txt 2>> $O/mpc.log
rm -f $O/mpc.ncu-rep
cat $O/mpc_summary.txt; grep -v " 0 \| n/a" $O/mpc_stalls.txt | sort -k2 -n -r | head -60; tail -3 $O/mpc.log
