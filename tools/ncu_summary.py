#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the JSON kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/ncu_rNN_name.json "free-text source line"
"""
import csv
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
]


def main():
    rep, out, src = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(text.splitlines()))
    head, units = rows[0], rows[1]
    kernels = []
    for r in rows[2:]:
        d = dict(zip(head, r))
        u = dict(zip(head, units))
        k = {"Kernel Name": d["Kernel Name"]}
        for key in KEYS:
            if key in d and d[key] != "":
                k[key] = {"value": d[key], "unit": u.get(key, "")}
        kernels.append(k)
    json.dump({"source": src, "tool": "ncu --set full --clock-control none; ncu -i <rep> --page raw --csv", "kernels": kernels},
              open(out, "w"), indent=1)
    print("wrote", out, len(kernels), "kernel(s)")


if __name__ == "__main__":
    main()
