#!/usr/bin/env python
"""Summarises an .ncu-rep (raw page) per kernel: the metrics DESIGN.md's roofline discussion uses.
Usage: python tools/ncu_summary.py report.ncu-rep [> profiles/rNN_xxx.txt]"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__occupancy_limit_registers", "occ limit regs (blocks)"),
    ("launch__occupancy_limit_shared_mem", "occ limit smem (blocks)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "IPC per SM (of 4)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp inst"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_fma.sum", "FMA pipe inst"), ("sm__inst_executed_pipe_fp64.sum", "FP64 pipe inst"),
    ("sm__inst_executed_pipe_alu.sum", "ALU pipe inst"), ("sm__inst_executed_pipe_lsu.sum", "LSU pipe inst"),
    ("sm__inst_executed_pipe_xu.sum", "XU pipe inst"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "FADD thread inst"),
    ("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "FMUL thread inst"),
    ("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "FFMA thread inst"),
    ("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "DADD thread inst"),
    ("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "DMUL thread inst"),
    ("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "DFMA thread inst"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__inst_executed_op_local_ld.sum", "local loads"), ("smsp__inst_executed_op_local_st.sum", "local stores"),
]
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
total = sum(float(r[hdr.index("gpu__time_duration.sum")]) for r in rows[2:])
print("report:", rep)
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    t = float(d["gpu__time_duration.sum"])
    print("=== {}  [{:.1f} % of the profiled launches]".format(d["Kernel Name"][:70], 100 * t / total))
    for k, label in KEYS:
        if k in d and d[k] != "":
            print("  {:34s} {:>18s} {}".format(label, d[k], u[k]))
    stalls = [(float(d[h]), h) for h in hdr if "issue_stalled" in h and h.endswith("_per_issue_active.ratio")
              and "not_issued" not in h and d[h] not in ("", "n/a")]
    stalls.sort(reverse=True)
    print("  top stalls (warps per issue):", ", ".join("{} {:.2f}".format(
        h.split("issue_stalled_")[1].split("_per_issue")[0], v) for v, h in stalls[:6]))
