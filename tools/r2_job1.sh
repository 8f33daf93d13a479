#!/bin/bash
# round 2, GPU job 1: tests, default bench (both arms), launch lists (default and --cache-control none), flop counts
O=gpurun_out/r2; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
(time python -m pytest tests -m gpu -x -q) > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
(time python bench.py --steps 20 --warmup 5) > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 600 $O/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_n1.json 2> $O/bench_ref_n1.err
bash tools/count_flops.sh
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_cfg2.csv python bench.py --steps 20 --warmup 3 --no-cpu --no-sub > $O/launches_cfg2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 300 --csv --log-file $O/launches_cfg2_nocacheflush.csv python bench.py --steps 20 --warmup 3 --no-cpu --no-sub > $O/launches_cfg2_nf.log 2>&1
python - <<'PY'
import json
for f in ("bench_n1", "bench_ref_n1"):
    try:
        d = json.loads(open(f"gpurun_out/r2/{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.4g" % d["value"], "ms/step %.4g" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "frac", (d.get("roofline") or {}).get("frac"))
        for k, v in (d.get("sub_records") or {}).items():
            print("  sub", k, v.get("error") or ("value %.4g ms %.4g e2e %.4g frac %s" % (v["value"], v["ms_per_step"], v["e2e"]["value"], v["roofline"].get("frac"))))
    except Exception as e:
        print(f, "FAILED", e)
PY
