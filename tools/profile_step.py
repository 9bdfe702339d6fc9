#!/usr/bin/env python
"""One profiled env step for ncu: warm up, then bracket a single step with cudaProfilerStart/Stop.
Usage (GPU box): ncu --profile-from-start off --set full --import-source on --clock-control none \
                     -o gpurun_out/step_<scene> python tools/profile_step.py <scene> [envs]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import scene_config
from safemotionsrisk_b200.vec_env import SafeMotionsVecEnv

scene = sys.argv[1] if len(sys.argv) > 1 else "ball"
envs = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
env = SafeMotionsVecEnv(num_envs=envs, config=scene_config(scene), seed=1)
env.reset()
for _ in range(25):
    env.step_random()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
env.step_random()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
env.close()
