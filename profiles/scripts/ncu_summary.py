#!/usr/bin/env python
"""Summarise an .ncu-rep: per profiled launch, the metrics the roofline argument needs.
usage: python profiles/scripts/ncu_summary.py <file.ncu-rep> [more metric substrings...]"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct", "smsp__warp_issue_stalled",
        "sm__pipe_tensor", "lts__t_sector_hit_rate.pct"]


def main():
    rep = sys.argv[1]
    keys = KEYS + sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== launch", r[hdr.index("ID")], r[hdr.index("Kernel Name")], "grid", r[hdr.index("Grid Size")], "block",
              r[hdr.index("Block Size")])
        for i, h in enumerate(hdr):
            if any(k in h for k in keys) and r[i] not in ("", "n/a"):
                print(f"   {h} = {r[i]} {units[i]}")


if __name__ == "__main__":
    main()
