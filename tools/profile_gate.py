#!/usr/bin/env python
"""One profiled risk gate (risk network + backup policy on the tensor cores + select) for ncu.
Usage (GPU box): ncu --profile-from-start off --set full --import-source on --clock-control none \
                     -o gpurun_out/gate python tools/profile_gate.py [scene]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import scene_config
from safemotionsrisk_b200.vec_env import SafeMotionsVecEnv

scene = sys.argv[1] if len(sys.argv) > 1 else "space_task_bm"
env = SafeMotionsVecEnv(num_envs=65536, config=scene_config(scene), seed=1)
env.load_networks()
env.reset()
for _ in range(10):
    env.step_gated(threshold=0.065)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
env.risk_gate(0.065)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
env.close()
