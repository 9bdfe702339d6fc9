#!/usr/bin/env python
"""Step time of experiment builds of the library against the default build.  Usage: lib_compare.py scene lib.so [lib.so ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
scene, libs = sys.argv[1], [None] + sys.argv[2:]
for lib in libs:
    env = dict(os.environ)
    if lib:
        env["SMENV_LIB"] = os.path.abspath(lib)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "step_time.py"), scene], env=env, capture_output=True, text=True)
    print(lib or "default", out.stdout.strip() or out.stderr[-300:], flush=True)
