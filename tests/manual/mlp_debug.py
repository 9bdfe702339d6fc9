import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from safemotionsrisk_b200 import ball_backup_config
from safemotionsrisk_b200.vec_env import SafeMotionsVecEnv
from oracle import mlp
env = SafeMotionsVecEnv(num_envs=1000, config=ball_backup_config(ball_machine_mode=True), fill_pools=(os.environ.get("FP","0")=="1"), auto_reset=False, seed=3)
env.load_networks()
w = np.load("safemotionsrisk_b200/assets/networks_ball.npz")
rng = np.random.default_rng(5)
obs = rng.uniform(-1, 1, (1000, 27)).astype(np.float32); act = rng.uniform(-1, 1, (1000, 7)).astype(np.float32)
for trial in ("full", "obs_only", "act_only", "zeros"):
    o, a = obs.copy(), act.copy()
    if trial == "obs_only": a[:] = 0
    if trial == "act_only": o[:] = 0
    if trial == "zeros": a[:] = 0; o[:] = 0
    r = env.mlp_forward(0, o, a, n_out=1).cpu().numpy()[:, 0]
    ref = mlp.risk_forward(w, o, a)
    e = np.abs(r - ref)
    print(trial, "max", e.max(), "mean", e.mean(), "bad rows", np.where(e > 1e-2)[0][:10], (e > 1e-2).sum(), r[:4], ref[:4])
for k in (0, 7, 8, 15, 16, 26, 27, 31, 32, 33):
    x = np.zeros((1000, 34), np.float32); x[:, k] = 1.0
    r = env.mlp_forward(0, x[:, :27], x[:, 27:], n_out=1).cpu().numpy()[:, 0]
    ref = mlp.risk_forward(w, x[:, :27], x[:, 27:])
    print("col", k, r[0], ref[0])
