#!/usr/bin/env python
"""Time of the risk gate (risk network + backup policy on the tensor cores + select) per step, 65536 envs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import scene_config
from safemotionsrisk_b200.vec_env import SafeMotionsVecEnv
scene = sys.argv[1] if len(sys.argv) > 1 else "space_bm"
env = SafeMotionsVecEnv(num_envs=65536, config=scene_config(scene), seed=1)
env.load_networks(); env.reset()
for _ in range(10): env.step_gated(threshold=0.065)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
tg = ts = 0.0
for _ in range(20):
    env._lib.smenv_random_actions(env._handle, env._buf, env._stream())
    e[0].record(); env.risk_gate(0.065); e[1].record()
    env._lib.smenv_step(env._handle, env._buf, 1, env._stream()); e[2].record()
    torch.cuda.synchronize()
    tg += e[0].elapsed_time(e[1]); ts += e[1].elapsed_time(e[2])
flops = 2 * 65536 * (30*512 + 512*256 + 256*128 + 128 + 23*256 + 256*128 + 128*7)
print(scene, "gate %.1f us/step (%.1f TFLOP/s dense-equivalent), env step %.1f us, risky fraction %.3f" % (
    1e3*tg/20, flops / (tg/20*1e-3) / 1e12, 1e3*ts/20, float(env.risky.float().mean())))
