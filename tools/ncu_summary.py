#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` dump: key roofline metrics + warp stall reasons per kernel.
usage: ncu -i X.ncu-rep --page raw --csv > raw.csv; python tools/ncu_summary.py raw.csv"""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_warps", "sm__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.max", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")][:90], " grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:78s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
                try:
                    stalls.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        print("  warp stall reasons (warps stalled per issue-active cycle):")
        for v, n in sorted(stalls, reverse=True)[:8]:
            print(f"    {n:40s} {v:7.2f}")


if __name__ == "__main__":
    main(sys.argv[1])
