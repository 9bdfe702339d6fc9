#!/usr/bin/env python
"""Prints the metrics of an `ncu --page raw --csv` export that the roofline discussion in DESIGN.md uses.
Usage: ncu -i prof.ncu-rep --page raw --csv | python tools/ncu_summary.py [regex]"""
import csv
import re
import sys

KEYS = r"gpu__time_duration.sum|launch__registers_per_thread|launch__occupancy_limit|launch__grid_size|launch__block_size|" \
       r"launch__waves|sm__warps_active.avg.pct|smsp__issue_active.avg.pct|sm__throughput.avg.pct|dram__bytes_(read|write).sum$|" \
       r"dram__throughput.avg.pct|sm__inst_executed_pipe_(fp64|fma|fmaheavy|alu|lsu|xu|uniform).*(sum|pct_of_peak_sustained_active)$|" \
       r"sm__pipe_fp64_cycles_active|smsp__inst_executed.sum$|smsp__thread_inst_executed_per_inst_executed.ratio|" \
       r"warp_issue_stalled.*_per_warp_active.pct|inst_executed_op_local|smsp__cycles_active.avg$|sm__cycles_elapsed.max|" \
       r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum|smsp__inst_executed_op_shared|lts__t_sector_hit_rate.pct"
pat = re.compile(sys.argv[1] if len(sys.argv) > 1 else KEYS)
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("=== kernel:", r[hdr.index("Kernel Name")][:60], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for h, u, v in zip(hdr, units, r):
        if pat.search(h):
            print("  {:90s} {:>18s} {}".format(h, v, u))
