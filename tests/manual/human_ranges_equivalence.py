#!/usr/bin/env python
"""Diagnostic: Human scene, device step in one range against the host-buffer step in c ranges."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from safemotionsrisk_b200 import abi, human_backup_config
from safemotionsrisk_b200.vec_env import SafeMotionsVecEnv
n = 1000
mk = lambda: SafeMotionsVecEnv(num_envs=n, config=human_backup_config(), seed=3, auto_reset=True)
ref = mk(); ref.reset(); ref.set_step_ranges(1)
envs = {c: mk() for c in (1, 3)}
for e in envs.values():
    e.reset()
rng = np.random.default_rng(11)
for step in range(40):
    act = rng.uniform(-1, 1, (n, 7)).astype(np.float32)
    ref.step(torch.from_numpy(act).cuda())
    torch.cuda.synchronize()
    for c, e in envs.items():
        obs, rew, done = e.step_host(act, chunks=c)
        names = ["obs", "rew", "done", "kin", "hkin", "hactions", "hstate", "hobs", "episode", "obst"]
        a = [obs, rew, done, e.kin.cpu().numpy(), e.hkin.cpu().numpy(), e.hactions.cpu().numpy(), e.hstate.cpu().numpy(), e.hobs.cpu().numpy(), e.episode.cpu().numpy(), e.obst.cpu().numpy()]
        b = [ref.obs.cpu().numpy(), ref.reward.cpu().numpy(), ref.done.cpu().numpy(), ref.kin.cpu().numpy(), ref.hkin.cpu().numpy(), ref.hactions.cpu().numpy(), ref.hstate.cpu().numpy(), ref.hobs.cpu().numpy(), ref.episode.cpu().numpy(), ref.obst.cpu().numpy()]
        bad = {nm: int((~np.isclose(x.reshape(n, -1).astype(np.float64), y.reshape(n, -1).astype(np.float64), rtol=0, atol=0, equal_nan=True)).any(1).sum()) for nm, x, y in zip(names, a, b)}
        bad = {k: v for k, v in bad.items() if v}
        if bad:
            rows = np.where((~np.isclose(a[4], b[4], rtol=0, atol=0)).any(1))[0][:5]
            print("step", step, "chunks", c, bad, "hkin rows", rows.tolist(), flush=True)
            if step > 6:
                sys.exit(0)
print("identical")
