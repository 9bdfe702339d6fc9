#!/usr/bin/env python
"""Diagnostic: the tensor-core MLP against a NumPy float32 forward pass for several layer widths."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ctypes as C
import numpy as np, torch
from oracle import mlp
from safemotionsrisk_b200 import cabi, space_backup_config
from safemotionsrisk_b200.vec_env import SafeMotionsVecEnv
env = SafeMotionsVecEnv(num_envs=1024, config=space_backup_config(ball_machine_mode=True), seed=1)
rng = np.random.default_rng(0)
x = rng.uniform(-1, 1, (1024, 30)).astype(np.float32)
for hid in ([512, 256, 128], [256, 128], [128, 64], [128, 128], [64, 64], [256, 64], [128, 64, 32], [64, 32], [32, 16], [512, 128], [192, 48], [96, 64]):
    dims = [30] + hid + [1]
    layers = [(rng.normal(0, 1.0 / np.sqrt(dims[i]), (dims[i], dims[i + 1])).astype(np.float32),
               rng.normal(0, 0.1, dims[i + 1]).astype(np.float32)) for i in range(len(dims) - 1)]
    flat = np.concatenate([np.concatenate([k.ravel(), b.ravel()]) for k, b in layers])
    d = np.array(dims, dtype=np.int32)
    rc = env._lib.smenv_mlp_load(env._handle, 0, len(hid), d.ctypes.data, 0, 0, flat.ctypes.data)
    if rc:
        print(hid, "load refused:", cabi.load().smenv_last_error().decode()[:120]); continue
    h = x
    for k, b in layers[:-1]:
        h = mlp.selu(h @ k + b)
    want = 1 / (1 + np.exp(-(h @ layers[-1][0] + layers[-1][1])))[:, 0]
    env._networks = True
    got = env.mlp_forward(0, x[:, :23], x[:, 23:])[:, 0].cpu().numpy()
    ex = env.mlp_forward_exact(0, x[:, :23], x[:, 23:])[:, 0].cpu().numpy()
    print(hid, "tensor cores max err {:.4f}  exact max err {:.2e}".format(np.abs(got - want).max(), np.abs(ex - want).max()), flush=True)
