#!/usr/bin/env python
"""Work of the GJK kernel per env-step (device counters): items by origin, iterations, support dots, iteration histogram.
Usage: gjk_counters.py [scene ...]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import scene_config
from safemotionsrisk_b200.vec_env import SafeMotionsVecEnv

for scene in sys.argv[1:] or ["space", "space_bm"]:
    env = SafeMotionsVecEnv(num_envs=65536, config=scene_config(scene), seed=1)
    env.reset()
    for _ in range(25):
        env.step_random()
    env.enable_counters(True)
    env.counters(reset=True)
    for _ in range(10):
        env.step_random()
    c = env.counters()
    n = c["env_steps"]
    print(scene, {k: round(v / n, 3) for k, v in c.items() if k not in ("gjk_iteration_histogram", "env_steps")},
          "hist(<=4,8,12,16,24,more)", [round(x / n, 3) for x in c["gjk_iteration_histogram"]], env.launch_config(), flush=True)
    env.close()
