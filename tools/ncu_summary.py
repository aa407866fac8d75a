#!/usr/bin/env python
"""Summarise an ncu report (raw page CSV) into the few numbers the roofline analysis needs.
    ncu -i X.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv [pattern ...]"""
import csv
import sys

DEFAULT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct",
           "gpu__dram_throughput", "lts__t_bytes.sum", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct",
           "sm__throughput.avg.pct", "sm__warps_active.avg.pct", "launch__registers_per_thread", "launch__occupancy",
           "launch__grid_size", "launch__block_size", "launch__waves", "sm__pipe_fp64_cycles_active", "smsp__issue_active.avg.pct",
           "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "issue_stalled", "smsp__inst_executed.sum",
           "sm__inst_executed_pipe", "lts__t_sectors_srcunit_tex_op", "l1tex__data_bank_conflicts", "smsp__warps_eligible",
           "dram__cycles_active", "sm__cycles_active.avg", "l1tex__t_bytes", "smsp__cycles_active.avg"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    pats = sys.argv[2:] or DEFAULT
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("=== %s  grid=%s block=%s" % (r[hdr.index("Kernel Name")][:80], r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
        for i, h in enumerate(hdr):
            if any(p in h for p in pats) and r[i] not in ("", "0", "n/a"):
                if "issue_stalled" in h and "not_issued" in h:
                    continue
                print("  %-95s %s %s" % (h, r[i], units[i]))


if __name__ == "__main__":
    main()
