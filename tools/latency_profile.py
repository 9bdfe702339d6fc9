#!/usr/bin/env python
"""Per-kernel times of one step against the env count (the latency floor of every kernel).
Usage: python tools/latency_profile.py [scene] [n ...]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import scene_config
from safemotionsrisk_b200.vec_env import SafeMotionsVecEnv

scene = sys.argv[1] if len(sys.argv) > 1 else "space"
for n in [int(a) for a in sys.argv[2:]] or [1, 1024, 8192, 32768, 65536]:
    env = SafeMotionsVecEnv(num_envs=n, config=scene_config(scene), seed=1)
    env.reset()
    for _ in range(25):
        env.step_random()
    env.kernel_timing(True)
    env.kernel_times(reset=True)
    for _ in range(30):
        env.step_random()
    t, k = env.kernel_times()
    print("{} n={:6d}: {}  sum {:.1f} us".format(scene, n, {a.replace("_kernel", ""): round(1e3 * b, 1) for a, b in t.items()},
                                                 1e3 * sum(t.values())), flush=True)
    env.close()
