#!/bin/bash
# One gpurun call that reproduces what the driver runs at round end, plus the profiles kept under profiles/:
#   gpurun --timeout 2400 -- 'bash tools/final_check.sh'
O=gpurun_out/final; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_cfg2.json 2> $O/bench_ref_cfg2.err
python bench.py > $O/bench_cfg2.json 2> $O/bench_cfg2.err
python bench.py --workload cfg3 --steps 30 --warmup 3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err
python bench.py --workload cfg4 --steps 500 --warmup 20 --no-cpu > $O/bench_cfg4.json 2> $O/bench_cfg4.err
python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu > $O/bench_cfg5.json 2> $O/bench_cfg5.err
python bench.py --workload loop2 --steps 300 --warmup 20 --no-cpu > $O/bench_loop2.json 2> $O/bench_loop2.err
python bench.py --workload loop4 --steps 200 --warmup 20 --no-cpu > $O/bench_loop4.json 2> $O/bench_loop4.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg2.csv python bench.py --steps 20 --warmup 3 --no-cpu > $O/launches_cfg2.log 2>&1
P=gpurun_out/p; mkdir -p $P
ncu --set full --clock-control none --import-source on -k regex:'cbfqp_kernel|odcbf_kernel' -s 0 -c 9 -f -o $P/qp python tools/prof_qp.py > $P/qp.log 2>&1
python tools/ncu_summary.py $P/qp.ncu-rep --title "qp: python tools/prof_qp.py" > $O/ncu_qp_summary.txt 2>> $P/qp.log
rm -f $P/qp.ncu-rep
bash tools/prof_mpc_ncu.sh 4096 > $O/ncu_mpc.log 2>&1
cp $P/mpc_summary.txt $O/ncu_mpc_summary.txt; cp $P/mpc_stalls.txt $O/ncu_mpc_stalls.txt
cat $O/pytest_gpu.log; tail -2 $O/smoke.log
for f in ref_cfg2 cfg2 cfg3 cfg4 cfg5 loop2 loop4; do python - <<PY
import json
try:
    d = json.loads(open("$O/bench_$f.json").read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print("$f", "value %.4g" % d["value"], "ms/step %.4g" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "frac", r.get("frac"), "launches", d.get("gpu_launches"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("$f", "FAILED", e)
PY
done
