#!/usr/bin/env python
"""Device time of the whole step (CUDA events, L2 flushed between steps like bench.py).  Usage: step_time.py [scene ...]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import scene_config
from safemotionsrisk_b200.vec_env import SafeMotionsVecEnv

for scene in sys.argv[1:] or ["ball", "space"]:
    env = SafeMotionsVecEnv(num_envs=65536, config=scene_config(scene), seed=1)
    env.reset()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(25):
        env.step_random()
    k = 100
    s = [torch.cuda.Event(enable_timing=True) for _ in range(k)]
    e = [torch.cuda.Event(enable_timing=True) for _ in range(k)]
    for i in range(k):
        flush.fill_(i & 255)
        s[i].record()
        env.step_random()
        e[i].record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in zip(s, e)) / k
    print("{:8s} {:.1f} us/step = {:.1f} M env-steps/s".format(scene, 1e3 * ms, 65536 / ms / 1e3), flush=True)
    env.close()
