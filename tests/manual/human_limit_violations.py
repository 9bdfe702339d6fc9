#!/usr/bin/env python
"""Diagnostic: how often and by how much does a human joint leave its position limits in a 65 536-env rollout, and at
which step of the nested env's episode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from safemotionsrisk_b200 import abi, human_backup_config
from safemotionsrisk_b200.vec_env import SafeMotionsVecEnv
n = 65536
env = SafeMotionsVecEnv(num_envs=n, config=human_backup_config(), seed=1, auto_reset=True)
sc = env.scene
env.reset()
hlo, hhi = (torch.tensor(x, device=env.device) for x in (sc.human_pos_lo, sc.human_pos_hi))
hstart, _ = env.human_pools()
print("pool: entries", hstart.shape, "outside limits:", int(((hstart[:, :8] < np.array(sc.human_pos_lo)) | (hstart[:, :8] > np.array(sc.human_pos_hi))).any(1).sum()))
for s in range(35):
    env.step_random()
    hq = env.hkin[:, 0:8]
    over = torch.clamp(hq - hhi, min=0) + torch.clamp(hlo - hq, min=0)
    bad = over.max(1).values > 1e-6
    if bool(bad.any()):
        idx = torch.nonzero(bad)[:, 0]
        steps = env.hstate[idx, abi.HS_STEPS].cpu().numpy()
        j = over[idx].argmax(1).cpu().numpy()
        print("step", s, "violating envs", int(bad.sum()), "max", float(over.max()), "episode steps", np.bincount(steps.astype(int))[:32].tolist(),
              "joints", np.bincount(j, minlength=8).tolist(), "braked", int((env.hstate[idx, abi.HS_BRAKED] != 0).sum()),
              "vel", env.hkin[idx[:3], 8:16].cpu().numpy().round(3).tolist()[:1])
env.close()
