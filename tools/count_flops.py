"""ncu launch csv (tools/count_flops.sh) -> profiles/r2_flops.json: FP64 flops per agent-solve of every MPC workload.

flops = 2 * DFMA + DADD + DMUL thread-level instructions (predicated-on) of the mpc_kernel launch / agents of that launch.
Also records, per launch: device time, FP64 pipe utilisation, warp instructions."""
import csv, json, sys

CASES = [("cfg3", "DynamicUnicycle2D", 4096), ("cfg5", "DynamicUnicycle2D", 2731), ("cfg5", "KinematicBicycle2D", 2731),
         ("cfg5", "Quad3D", 2730)]
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
per = {}
for r in rows[1:]:
    if len(r) < len(hdr) or "mpc_kernel" not in r[col["Kernel Name"]]:
        continue
    per.setdefault(r[col["ID"]], {})[r[col["Metric Name"]]] = float(r[col["Metric Value"]].replace(",", ""))
launches = [per[k] for k in sorted(per, key=int)]
assert len(launches) == len(CASES), (len(launches), "mpc_kernel launches; expected", len(CASES))
out = {"_how": "ncu --metrics smsp__sass_thread_inst_executed_op_{dfma,dadd,dmul}_pred_on.sum on ONE solve per workload "
               "(tools/count_flops.sh -> tools/prof_flops.py); flops = 2 DFMA + DADD + DMUL per agent", "cfg5": {}, "_detail": {}}
for (key, model, n), m in zip(CASES, launches):
    fl = 2 * m["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"] + m["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"] + \
        m["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"]
    if key == "cfg3":
        out["cfg3"] = fl / n
    else:
        out["cfg5"][model] = fl / n
    out["_detail"][f"{key}:{model}"] = {"agents": n, "flops": fl, "gpu_time_us": m.get("gpu__time_duration.sum", 0) / 1e3,
                                        "fp64_pipe_pct": m.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                                        "warp_inst": m.get("smsp__inst_executed.sum"),
                                        "tflops_under_ncu": fl / (m.get("gpu__time_duration.sum", 1) * 1e-9) / 1e12}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
