#!/bin/bash
# round 2, GPU job 2 (every stage under its own timeout): default bench (both arms), flop counts, launch lists, TMA A/B, MPC variants, tests
O=gpurun_out/r2; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt; nproc >> $O/smi.txt
(time timeout 900 python bench.py --steps 20 --warmup 5) > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 400 $O/bench_n1.err
timeout 400 python tools/sweep_tma.py > $O/sweep_tma.txt 2>&1; tail -20 $O/sweep_tma.txt
timeout 600 bash tools/count_flops.sh
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_cfg2.csv python bench.py --steps 20 --warmup 3 --no-cpu --no-sub > $O/launches_cfg2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 300 --csv --log-file $O/launches_cfg2_nocacheflush.csv python bench.py --steps 20 --warmup 3 --no-cpu --no-sub > $O/launches_cfg2_nf.log 2>&1
for v in "" _l16; do SCB_LIB=$PWD/safe_control_b200/libscb$v.so timeout 300 python tools/mpc_variants.py cfg3 du5 kb5 q5 >> $O/mpc_variants.txt 2>&1; done; cat $O/mpc_variants.txt
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_n1.json 2> $O/bench_ref_n1.err
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $O/pytest_gpu.log 2>&1; tail -15 $O/pytest_gpu.log
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
