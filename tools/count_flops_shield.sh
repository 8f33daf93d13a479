#!/bin/bash
# FP64 flop count of one gatekeeper step (both launches) and one MPS step at 65 536 agents -> gpurun_out/r2/r2_flops_shield.json
O=gpurun_out/r2; mkdir -p $O
ncu --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum \
    --clock-control none -k regex:'shield_step_kernel' -c 6 --csv --log-file $O/flops_shield.csv python tools/prof_shield.py 65536 > $O/flops_shield.log 2>&1
python - <<'PY'
import csv, json
rows = [r for r in csv.reader(open("gpurun_out/r2/flops_shield.csv")) if len(r) > 6]
hdr = rows[0]; ci = {h: i for i, h in enumerate(hdr)}
launch = {}
for r in rows[1:]:
    if not r[0].isdigit(): continue
    launch.setdefault(int(r[0]), {"name": r[ci["Kernel Name"]].split("(")[0]})[r[ci["Metric Name"]]] = float(r[ci["Metric Value"]].replace(",", ""))
ids = sorted(launch)
fl = lambda m: 2 * m["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"] + m["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"] + m["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"]
N = 65536
# launch order of tools/prof_shield.py: gatekeeper step 1 (phase 1, phase 2), gatekeeper step 2 (phase 1, phase 2), MPS step 1, MPS step 2
gk = [launch[i] for i in ids[2:4]]; mps = [launch[ids[5]]]
out = {"_how": "ncu --metrics smsp__sass_thread_inst_executed_op_{dfma,dadd,dmul}_pred_on.sum on the SECOND control step of 65536 agents (tools/count_flops_shield.sh -> tools/prof_shield.py); flops = 2 DFMA + DADD + DMUL",
       "agents": N,
       "gatekeeper": {"flops_per_agent_step": sum(fl(m) for m in gk) / N, "launches": [{"kernel": m["name"], "flops": fl(m), "gpu_time_ns": m["gpu__time_duration.sum"], "fp64_pipe_pct": m["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"], "warp_inst": m["smsp__inst_executed.sum"]} for m in gk]},
       "mps": {"flops_per_agent_step": sum(fl(m) for m in mps) / N, "launches": [{"kernel": m["name"], "flops": fl(m), "gpu_time_ns": m["gpu__time_duration.sum"], "fp64_pipe_pct": m["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"], "warp_inst": m["smsp__inst_executed.sum"]} for m in mps]}}
json.dump(out, open("gpurun_out/r2/r2_flops_shield.json", "w"), indent=1)
print(json.dumps(out, indent=1))
PY
