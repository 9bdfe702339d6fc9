#!/usr/bin/env python
"""Step time against the number of env ranges the device step runs side by side (SMENV_STEP_RANGES).  Usage: ranges_sweep.py scene ..."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for scene in sys.argv[1:] or ["space_bm"]:
    for r in (1, 2, 3, 4):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "step_time.py"), scene],
                             env=dict(os.environ, SMENV_STEP_RANGES=str(r)), capture_output=True, text=True)
        print("ranges", r, out.stdout.strip() or out.stderr[-300:], flush=True)
