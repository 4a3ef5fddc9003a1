#!/usr/bin/env python
"""Turn an `ncu --set full` report into the markdown table kept under profiles/.

usage: tools/ncu_summary.py report.ncu-rep "header line(s)" > profiles/rN_name_ncu.md
"""
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "smsp__inst_executed.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
    "l1tex__data_bank_conflicts_pipe_lsu.sum",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
    "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_lg_throttle",
    "smsp__pcsamp_warps_issue_stalled_sleeping", "smsp__pcsamp_warps_issue_stalled_wait",
    "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_selected",
    "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_no_instructions",
]


def main():
    rep, header = sys.argv[1], sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True,
                         check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    names, units, vals = rows[0], rows[1], rows[2]
    for h in header:
        print("# " + h)
    print("\n| metric | unit | value |\n|---|---|---|")
    col = {n: i for i, n in enumerate(names)}
    print(f"| kernel | | `{vals[col['Kernel Name']]}` |")
    for k in KEEP:
        if k in col:
            print(f"| {k} | {units[col[k]]} | {vals[col[k]]} |")


if __name__ == "__main__":
    main()
