#!/usr/bin/env python
"""Diagnostic: the `mixed` env of test_host_step_in_ranges_equals_the_device_step for the Human scene."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from safemotionsrisk_b200 import human_backup_config
from safemotionsrisk_b200.vec_env import SafeMotionsVecEnv
n = 1000
mk = lambda: SafeMotionsVecEnv(num_envs=n, config=human_backup_config(), seed=3, auto_reset=True)
ref = mk(); ref.reset(); ref.set_step_ranges(1)
mixed = mk(); mixed.reset(); mixed.set_step_ranges(3)
rng = np.random.default_rng(11)
for step in range(45):
    act = rng.uniform(-1, 1, (n, 7)).astype(np.float32)
    ref.step(torch.from_numpy(act).cuda())
    torch.cuda.synchronize()
    c = 1 + step % 5
    if step % 7 == 3:
        mixed.step(torch.from_numpy(act).cuda()); how = "device x3"
    else:
        mixed.step_host(act, chunks=c); how = "host x{}".format(c)
    torch.cuda.synchronize()
    names = ["obs", "reward", "done", "kin", "hkin", "hactions", "hstate", "hobs", "episode", "obst", "actions"]
    bad = {}
    for nm in names:
        a, b = getattr(mixed, nm).cpu().numpy().reshape(n, -1).astype(np.float64), getattr(ref, nm).cpu().numpy().reshape(n, -1).astype(np.float64)
        rows = np.where((~np.isclose(a, b, rtol=0, atol=0, equal_nan=True)).any(1))[0]
        if len(rows):
            bad[nm] = (len(rows), rows[:6].tolist())
    if bad:
        print("step", step, how, bad, flush=True)
        break
else:
    print("identical")
