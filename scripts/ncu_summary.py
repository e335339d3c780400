#!/usr/bin/env python
"""Summarise an .ncu-rep (one or more launches) into the handful of numbers the
roofline discussion needs.  usage: ncu_summary.py report.ncu-rep [--all]"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum",
    "sm__inst_executed_pipe_cbu.sum", "sm__inst_executed_pipe_adu.sum",
    "smsp__inst_executed_op_shfl.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__warps_active.avg.per_cycle_active",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("== kernel:", d.get("Kernel Name", "?"), "| id", d.get("ID"))
        for h, u in zip(hdr, units):
            v = d[h]
            if any(h == k or (h.startswith(k) and k.endswith("sum") is False and h == k) for k in KEYS):
                print(f"  {h:78s} {v:>16s} {u}")
        stalls = []
        for h in hdr:
            if h.startswith(STALL) and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(d[h]), h[len(STALL):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("  stall reasons (warps stalled per issue-active cycle):",
              ", ".join(f"{n} {v:.2f}" for v, n in stalls[:9]))
        if "--all" not in sys.argv:
            break


if __name__ == "__main__":
    main()
