#!/usr/bin/env python
"""Key numbers of an `ncu --page raw --csv` export: duration, pipe utilisation, occupancy, stall reasons, DRAM bytes.

usage: python profiles/raw_summary.py <raw_page.csv>
"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__ctas_launched.sum", "launch__grid_size", "launch__block_size"]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("=====", r[hdr.index("Kernel Name")][:60])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:<70s} {r[hdr.index(k)]} {units[hdr.index(k)]}")
        items = []
        for i, hn in enumerate(hdr):
            if "pcsamp_warps_issue_stalled" in hn and "not_issued" not in hn:
                try:
                    items.append((float(r[i]), hn.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(x for x, _ in items) or 1.0
        print("  stall samples:", ", ".join(f"{n} {100 * x / tot:.1f}%" for x, n in sorted(items, reverse=True)[:9]))


if __name__ == "__main__":
    main(sys.argv[1])
