#!/usr/bin/env python3
"""Summarise an .ncu-rep (read on the CPU box): python profiles/ncu_summary.py file.ncu-rep [kernel-regex]"""
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size", "sm__inst_executed.sum",
    "smsp__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fma.sum",
    "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
    "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum", "smsp__inst_executed_op_global_red.sum",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if len(sys.argv) > 2 and not re.search(sys.argv[2], name):
            continue
        print("== kernel:", name, "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
        for i, h in enumerate(hdr):
            short = h.split(".", 2)[-1] if h.count(".") > 2 and h.split(".")[1].startswith("Triage") else h
            if h in KEYS or "issue_stalled" in h and h.endswith("_per_warp_active.pct"):
                print("  %-90s %12s %s" % (h, r[i], units[i]))


if __name__ == "__main__":
    main()
