#!/usr/bin/env python
"""Device time of the step and of the GJK kernel for several shared-memory configurations of the GJK kernel
(direction tables for hulls of at least SMENV_LUT_MIN_VERTS vertices; 256 threads x 2 CTAs or 512 x 1 per SM).
Usage: gjk_config_sweep.py scene [scene ...]   (each configuration runs in a fresh process: the env var is read once)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, torch
sys.path.insert(0, %r)
from bench import scene_config
from safemotionsrisk_b200.vec_env import SafeMotionsVecEnv
env = SafeMotionsVecEnv(num_envs=65536, config=scene_config(sys.argv[1]), seed=1)
env.reset()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(25):
    env.step_random()
k = 60
s = [torch.cuda.Event(enable_timing=True) for _ in range(k)]
e = [torch.cuda.Event(enable_timing=True) for _ in range(k)]
for i in range(k):
    flush.fill_(i & 255); s[i].record(); env.step_random(); e[i].record()
torch.cuda.synchronize()
ms = sum(a.elapsed_time(b) for a, b in zip(s, e)) / k
env.kernel_timing(True); env.kernel_times(reset=True)
for i in range(30):
    flush.fill_(i & 255); env.step_random()
kt, _ = env.kernel_times()
print(json.dumps(dict(scene=sys.argv[1], us_per_step=1e3 * ms, gjk_us=1e3 * kt["gjk_kernel"], launch=env.launch_config())))
''' % ROOT

# SMENV_SWEEP: JSON list of env-var overrides, one process each (default: CTA geometry of the GJK kernel)
CONFIGS = json.loads(os.environ.get("SMENV_SWEEP", "null")) or [
    {}, {"SMENV_GJK_THREADS": "768"}, {"SMENV_GJK_THREADS": "1024"}, {"SMENV_GJK_256": "1"}]
for scene in sys.argv[1:] or ["space_bm"]:
    for env_over in CONFIGS:
        out = subprocess.run([sys.executable, "-c", CHILD, scene], env=dict(os.environ, **env_over), capture_output=True, text=True)
        line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:]
        print(env_over, line, flush=True)
