#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small JSON/markdown-friendly dict of the counters DESIGN.md cites.
Usage: python profiles/summarize.py gpurun_out/prof.ncu-rep [kernel-regex] > profiles/<name>.json"""
import csv
import io
import json
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
    "smsp__sass_average_branch_targets_threads_uniform.pct",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        name = d.get("Kernel Name", "?")
        if len(sys.argv) > 2 and not re.search(sys.argv[2], name):
            continue
        rec = {"kernel": name}
        for h, u, v in zip(hdr, units, vals):
            if h in KEYS:
                try:
                    rec[h] = {"value": float(v.replace(",", "")), "unit": u}
                except ValueError:
                    rec[h] = {"value": v, "unit": u}
        res.append(rec)
    json.dump(res, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
