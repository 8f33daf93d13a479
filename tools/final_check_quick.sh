#!/bin/bash
# 1 GPU: the driver's round-end sequence (GPU tests, smoke, reference arm, default bench) + sanitizer over every kernel family
#   gpurun --timeout 1500 -- 'bash tools/final_check_quick.sh'
O=gpurun_out/final; mkdir -p $O
(time timeout 1700 python -m pytest tests -m gpu -x -q) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
(time timeout 400 python bench.py --impl reference --steps 20 --warmup 5) > $O/bench_ref_n1.json 2> $O/bench_ref_n1.err; tail -3 $O/bench_ref_n1.err
(time timeout 900 python bench.py --steps 20 --warmup 5) > $O/bench_n1.json 2> $O/bench_n1.err; tail -3 $O/bench_n1.err
timeout 1500 bash tools/sanitize.sh > $O/sanitize.log 2>&1; tail -12 $O/sanitize.log
cp gpurun_out/r2/sanitizer_memcheck.log $O/; cp gpurun_out/r2/sanitizer_racecheck.log $O/
python - <<'PY'
import json
for f in ("bench_ref_n1", "bench_n1"):
    try:
        d = json.loads(open(f"gpurun_out/final/{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.4g" % d["value"], "ms/step %.4g" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "frac", (d.get("roofline") or {}).get("frac"))
        for k, v in (d.get("sub_records") or {}).items():
            print("  sub", k, v.get("error") or ("value %.4g ms %.4g e2e %.4g frac %s" % (v["value"], v["ms_per_step"], v["e2e"]["value"], v["roofline"].get("frac"))))
    except Exception as e:
        print(f, "FAILED", e)
PY
