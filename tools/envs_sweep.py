#!/usr/bin/env python
"""Device time of the step against the number of envs per GPU (4 096 ... 262 144), L2 flushed between steps like bench.py.
Usage: python tools/envs_sweep.py [scene ...]   -> one line per (scene, envs): us per step, env-steps/s"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import scene_config
from safemotionsrisk_b200.vec_env import SafeMotionsVecEnv

for scene in sys.argv[1:] or ["human", "space"]:
    for n in (4096, 8192, 16384, 32768, 65536, 131072, 262144):
        env = SafeMotionsVecEnv(num_envs=n, config=scene_config(scene), seed=1)
        env.reset()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        for _ in range(15):
            env.step_random()
        k = 40
        s = [torch.cuda.Event(enable_timing=True) for _ in range(k)]
        e = [torch.cuda.Event(enable_timing=True) for _ in range(k)]
        for i in range(k):
            flush.fill_(i & 255)
            s[i].record()
            env.step_random()
            e[i].record()
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in zip(s, e)) / k
        print("{:8s} envs {:7d}  {:8.1f} us/step  {:7.2f} M env-steps/s  ranges {}".format(
            scene, n, 1e3 * ms, n / ms / 1e3, env.launch_config()["step_ranges"]), flush=True)
        env.close()
        del env, flush
        torch.cuda.empty_cache()
