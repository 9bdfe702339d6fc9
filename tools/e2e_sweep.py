#!/usr/bin/env python
"""Host-buffer step time against the number of env ranges (smenv_step_host).  Usage: python tools/e2e_sweep.py [scene ...]"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import scene_config
from safemotionsrisk_b200.vec_env import SafeMotionsVecEnv

for scene in sys.argv[1:] or ["ball", "space"]:
    env = SafeMotionsVecEnv(num_envs=65536, config=scene_config(scene), seed=1)
    env.reset()
    acts = np.random.default_rng(0).uniform(-1, 1, (65536, 7)).astype(np.float32)
    for _ in range(20):
        env.step_random()
    for chunks in (1, 2, 3, 4, 6, 8):
        for _ in range(5):
            env.step_host(acts, chunks=chunks)
        t0 = time.perf_counter()
        for _ in range(50):
            env.step_host(acts, chunks=chunks)
        dt = (time.perf_counter() - t0) / 50
        t0 = time.perf_counter()
        for _ in range(50):
            env.step_host(None, chunks=chunks)   # actions already in the pinned buffer
        dt2 = (time.perf_counter() - t0) / 50
        print("{:6s} chunks {}: {:.3f} ms/step = {:.1f} M env-steps/s   (pinned actions in place: {:.3f} ms)".format(
            scene, chunks, 1e3 * dt, 65536 / dt / 1e6, 1e3 * dt2), flush=True)
    env.close()
