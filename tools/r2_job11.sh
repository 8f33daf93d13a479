#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
timeout 300 bash tools/count_flops_backup.sh 2>&1 | tail -30
cp $O/r2_flops_backup.json profiles/r2_flops_backup.json
(time timeout 600 python bench.py --workload backup --steps 20 --warmup 3) > $O/bench_backup.json 2> $O/bench_backup.err
tail -c 400 $O/bench_backup.err; cat $O/bench_backup.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.4g ms %.4g e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['roofline']['achieved'], d['roofline']['frac'], d['config']['single_agent_latency_ms'], d.get('cpu_baseline'))"
