#!/bin/bash
# e2e of cfg2 through the host-pointer ABI with different kernel geometries (zero-copy reads over PCIe)
O=gpurun_out/r2; mkdir -p $O
for L in "" 32 8 4; do
  echo "== SCB_QP_LANES=${L:-auto}"
  SCB_QP_LANES=$L timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-sub --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); e = d['e2e']
print('value %.4g ms %.4g | e2e %.4g' % (d['value'], d['ms_per_step'], e['value']), {k: (round(v['value']/1e6, 2) if isinstance(v, dict) and 'value' in v else v) for k, v in e.items() if k not in ('value', 'unit', 'how')})"
done 2>&1 | tee $O/e2e_lanes.txt
