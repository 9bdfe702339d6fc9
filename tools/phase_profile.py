#!/usr/bin/env python
"""Per-phase cycle shares of the geometry kernel from the counting build (clock64 deltas summed over warps).
Usage (GPU box): python tools/phase_profile.py [scene ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import scene_config
from safemotionsrisk_b200.vec_env import SafeMotionsVecEnv
NAMES = ["load", "contact broad", "contact narrow", "end-pose FK", "static+self dist", "moving dist", "reward/out", "obs/reset"]
for scene in (sys.argv[1:] or ["ball", "space"]):
    env = SafeMotionsVecEnv(num_envs=65536, config=scene_config(scene), seed=1)
    env.reset()
    for _ in range(25): env.step_random()
    env.enable_counters(True); env.counters(reset=True)
    for _ in range(5): env.step_random()
    c = env.counters(); n = c["env_steps"]
    tot = sum(c["phase_cycles"])
    print(scene, "per env-step: gjk calls %.2f iters %.2f dots %.0f flagged sub-steps %.2f cycles/warp %.0f" % (
        c["gjk_calls"]/n, c["gjk_iters"]/n, c["support_dots"]/n, c["flagged_substeps"]/n, tot/n))
    for name, cyc in zip(NAMES, c["phase_cycles"]):
        print("   %-18s %8.0f cycles  %5.1f %%" % (name, cyc/n, 100*cyc/tot))
    env.close()
