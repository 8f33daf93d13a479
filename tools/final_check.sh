#!/bin/bash
# One gpurun call (1 GPU) that reproduces what the driver runs at round end, plus the profiles kept under profiles/r2_*:
#   gpurun --timeout 3000 -- 'bash tools/final_check.sh'
# Multi-GPU lines:  gpurun --gpus N --timeout 900 -- 'bash tools/r2_jobN.sh N'     (N = 2, 4, 8)
O=gpurun_out/final; mkdir -p $O
(time timeout 1700 python -m pytest tests -m gpu -x -q) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref_n1.json 2> $O/bench_ref_n1.err
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 300 python bench.py --workload loop2 --steps 300 --warmup 20 --no-cpu > $O/bench_loop2.json 2> $O/bench_loop2.err
timeout 300 python bench.py --workload loop4 --steps 200 --warmup 20 --no-cpu > $O/bench_loop4.json 2> $O/bench_loop4.err
timeout 300 python tools/sweep_tma.py > $O/sweep_tma.txt 2>&1
timeout 600 bash tools/count_flops.sh
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_cfg2.csv python bench.py --steps 20 --warmup 3 --no-cpu --no-sub > $O/launches_cfg2.log 2>&1
timeout 1200 bash tools/prof_r2.sh > $O/prof.log 2>&1
timeout 1500 bash tools/sanitize.sh > $O/sanitize.log 2>&1
python - <<'PY'
import json
for f in ("bench_ref_n1", "bench_n1", "bench_loop2", "bench_loop4"):
    try:
        d = json.loads(open(f"gpurun_out/final/{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.4g" % d["value"], "ms/step %.4g" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "frac", (d.get("roofline") or {}).get("frac"))
        for k, v in (d.get("sub_records") or {}).items():
            print("  sub", k, v.get("error") or ("value %.4g ms %.4g e2e %.4g frac %s" % (v["value"], v["ms_per_step"], v["e2e"]["value"], v["roofline"].get("frac"))))
    except Exception as e:
        print(f, "FAILED", e)
PY
