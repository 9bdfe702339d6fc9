#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel from an .ncu-rep captured with --import-source on.
Usage: python tools/ncu_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern,
                      "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = [i for i, r in enumerate(rows) if len(r) > 5 and r[0] == "Line No"][0]
hdr = rows[hi]
ie, it, isamp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
cur, agg = None, {}
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 10 and r[0] not in ("", "Line No"):
        try:
            key = (cur, int(r[0]), r[1].strip()[:100])
            agg[key] = (int(r[ie]), int(r[it]), int(r[isamp]))
        except ValueError:
            pass
tot = sum(v[0] for v in agg.values())
tots = sum(v[2] for v in agg.values())
print("kernel", kern, "warp instructions", tot, "samples", tots)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% samples  %4.1f thr/inst  %s:%d  %s" % (
        100 * v[0] / tot, 100 * v[2] / max(tots, 1), v[1] / max(v[0], 1), k[0], k[1], k[2]))
