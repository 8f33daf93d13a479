#!/usr/bin/env python
"""Summarise a .ncu-rep (from `ncu --set full`) into the text files kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--title "..."] > profiles/rN_ncu_<kernel>_summary.txt

Reads the report with `ncu -i <rep> --page raw --csv` (works on the GPU-less build box) and prints, per launch,
the metrics the roofline discussion in DESIGN.md uses, plus the top stall reasons.
"""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sass__inst_executed_local_loads",
    "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    title = sys.argv[sys.argv.index("--title") + 1] if "--title" in sys.argv else rep
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {title}")
    print("# source: ncu --set full --clock-control none --import-source on (binary report not committed); "
          "summarised by tools/ncu_summary.py\n")
    for r in data:
        name = r[col["Kernel Name"]]
        grid, block = r[col.get("Grid Size", 0)], r[col.get("Block Size", 0)]
        print(f"## {name[:110]}   grid {grid} block {block}")
        for k in KEEP:
            if k in col and r[col[k]] != "":
                print(f"{k:<86s} {r[col[k]]:>16s} {units[col[k]]}")
        stalls = []
        for h, i in col.items():
            if h.startswith(STALL) and h.endswith("_per_warp_active.pct") and r[i] not in ("", "n/a"):
                try:
                    stalls.append((float(r[i].replace(",", "")), h[len(STALL):-len("_per_warp_active.pct")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        if stalls:
            print("top stall reasons (% of warp-active): " + ", ".join(f"{n} {v:.1f}" for v, n in stalls[:6]))
        print()


if __name__ == "__main__":
    main()
