#!/bin/bash
# FP64 flop count of one Backup-CBF solve (rollout + QP launches) at 65 536 agents -> gpurun_out/r2/r2_flops_backup.json
O=gpurun_out/r2; mkdir -p $O
ncu --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum \
    --clock-control none -k regex:'backup_rollout_kernel|backup_qp_kernel' -c 2 --csv --log-file $O/flops_backup.csv python tools/prof_backup.py 65536 > $O/flops_backup.log 2>&1
python - <<'PY'
import csv, json
rows = [r for r in csv.reader(open("gpurun_out/r2/flops_backup.csv")) if len(r) > 6]
hdr = rows[0]; ci = {h: i for i, h in enumerate(hdr)}
acc = {}
for r in rows[1:]:
    if not r[0].isdigit(): continue
    k = r[ci["Kernel Name"]].split("(")[0]
    acc.setdefault(k, {})[r[ci["Metric Name"]]] = float(r[ci["Metric Value"]].replace(",", ""))
N = 65536
out = {"_how": "ncu --metrics smsp__sass_thread_inst_executed_op_{dfma,dadd,dmul}_pred_on.sum on ONE solve of 65536 agents (tools/count_flops_backup.sh -> tools/prof_backup.py); flops = 2 DFMA + DADD + DMUL", "agents": N, "kernels": {}}
tot = 0.0
for k, m in acc.items():
    fl = 2 * m["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"] + m["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"] + m["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"]
    tot += fl
    out["kernels"][k] = {"flops": fl, "gpu_time_us": m["gpu__time_duration.sum"] / (1e3 if m["gpu__time_duration.sum"] > 1e5 else 1), "fp64_pipe_pct": m["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"], "warp_inst": m["smsp__inst_executed.sum"]}
out["flops_per_agent"] = tot / N
json.dump(out, open("gpurun_out/r2/r2_flops_backup.json", "w"), indent=1)
print(json.dumps(out, indent=1))
PY
