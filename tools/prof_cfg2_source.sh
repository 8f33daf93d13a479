#!/bin/bash
# per-instruction stall samples of the cfg2 kernel (warp per QP) -> gpurun_out/r2/ncu_cfg2_source.csv (SASS + samples)
O=gpurun_out/r2; P=gpurun_out/p; mkdir -p $O $P
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'cbfqp_kernel' -s 1 -c 1 -f -o $P/cfg2 python tools/prof_qp.py > $P/cfg2.log 2>&1
ncu -i $P/cfg2.ncu-rep --page source --csv > $O/ncu_cfg2_source.csv 2>> $P/cfg2.log
ncu -i $P/cfg2.ncu-rep --page raw --csv | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin)); h, u, d = rows[0], rows[1], rows[2]
for i, n in enumerate(h):
    if 'stall' in n and 'per_warp_active' in n or 'inst_executed' in n and 'sum' in n or 'time_duration' in n:
        print(f'{n:100s} {d[i]:>14s} {u[i]}')
" > $O/ncu_cfg2_stalls.txt
rm -f $P/cfg2.ncu-rep; wc -l $O/ncu_cfg2_source.csv; head -3 $O/ncu_cfg2_source.csv | cut -c1-600
