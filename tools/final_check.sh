#!/bin/bash
# One gpurun call (1 GPU) that reproduces what the driver runs at round end, plus the profiles kept under profiles/r2_*:
#   gpurun --timeout 3400 -- 'bash tools/final_check.sh'
# tools/final_check_quick.sh is the first half alone (GPU suite, smoke, reference arm, default bench, sanitizer: ~5.5 min).
# Multi-GPU lines:  gpurun --gpus N --timeout 900 -- 'bash tools/r2_jobN.sh N'     (N = 2, 4, 8)
bash tools/final_check_quick.sh
O=gpurun_out/final; R=gpurun_out/r2; mkdir -p $O $R
timeout 300 python bench.py --workload loop2 --steps 300 --warmup 20 --no-cpu > $O/bench_loop2.json 2> $O/bench_loop2.err
timeout 300 python bench.py --workload loop4 --steps 200 --warmup 20 --no-cpu > $O/bench_loop4.json 2> $O/bench_loop4.err
for w in backup gatekeeper mps; do timeout 300 python bench.py --workload $w --steps 20 --warmup 3 > $O/bench_$w.json 2> $O/bench_$w.err; done
# launch-geometry A/Bs
timeout 300 python tools/sweep_tma.py > $O/sweep_tma.txt 2>&1
timeout 300 python tools/mpc_variants.py cfg3 du5 kb5 q5 > $O/mpc_variants.txt 2>&1
timeout 300 python tools/time_backup.py 65536 10 > $O/backup_timing.txt 2>&1
timeout 300 python tools/time_shield.py 65536 > $O/shield_timing.txt 2>&1
# counted flops (numerators of the FP64 rooflines) and ncu captures
timeout 600 bash tools/count_flops.sh; timeout 300 bash tools/count_flops_backup.sh; timeout 300 bash tools/count_flops_shield.sh
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_cfg2.csv python bench.py --steps 20 --warmup 3 --no-cpu --no-sub > $O/launches_cfg2.log 2>&1
timeout 1200 bash tools/prof_r2.sh > $O/prof.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'backup' -s 0 -c 3 -f -o gpurun_out/p/bk python tools/prof_backup.py 65536 > gpurun_out/p/bk.log 2>&1
python tools/ncu_summary.py gpurun_out/p/bk.ncu-rep --title "backup-cbf: python tools/prof_backup.py 65536" > $R/ncu_backup_summary.txt; rm -f gpurun_out/p/bk.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'shield_step' -s 2 -c 2 -f -o gpurun_out/p/sh python tools/prof_shield.py 65536 > gpurun_out/p/sh.log 2>&1
python tools/ncu_summary.py gpurun_out/p/sh.ncu-rep --title "gatekeeper: python tools/prof_shield.py 65536 (second step, both launches)" > $R/ncu_shield_summary.txt; rm -f gpurun_out/p/sh.ncu-rep
python - <<'PY'
import json
for f in ("bench_loop2", "bench_loop4", "bench_backup", "bench_gatekeeper", "bench_mps"):
    try:
        d = json.loads(open(f"gpurun_out/final/{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.4g" % d["value"], "ms/step %.4g" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "frac", (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "FAILED", e)
PY
