#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
(time timeout 1200 python -m pytest tests/test_gpu_shield.py -x -q -m gpu) > $O/pytest_gpu_shield.log 2>&1
tail -5 $O/pytest_gpu_shield.log
timeout 300 python tools/time_shield.py 65536 2>&1 | tee $O/shield_timing.txt
for w in gatekeeper mps; do
  (time timeout 600 python bench.py --workload $w --steps 20 --warmup 3) > $O/bench_$w.json 2> $O/bench_$w.err
  tail -c 300 $O/bench_$w.err; python -c "
import sys, json
d = json.loads(open('$O/bench_$w.json').read().strip().splitlines()[-1])
print('$w value %.4g ms %.4g e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['config']['outcome'], d.get('cpu_baseline'))"
done
