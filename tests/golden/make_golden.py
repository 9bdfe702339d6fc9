#!/usr/bin/env python
"""Generates tests/golden/*.npz from the CPU oracle.

The reference cannot run in this image (pybullet / klimits / gym absent) and ships no golden vectors,
so these fixtures pin the ORACLE (regression vectors), not the reference: inputs (start states, actions, ball
launches) and the oracle's outputs per step.  `tests/test_needs_pybullet.py` documents how to re-record them from the
real env on a box that has pybullet + klimits.
Usage: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from safemotionsrisk_b200 import ball_backup_config, human_backup_config, space_backup_config  # noqa: E402
from safemotionsrisk_b200.scene import Scene  # noqa: E402

N, STEPS = 24, 20


def synth_ball(rng, n):
    """Ball launches aimed at the workspace from the r = 2.5 m sphere (README.md:81), without collision checks."""
    out = np.zeros((n, 12))
    for i in range(n):
        h = rng.uniform(-0.5, 0.5)
        ang = rng.uniform(0, 6.2831)
        r = np.sqrt(2.5 ** 2 - h ** 2)
        rel = np.array([r * np.cos(ang), r * np.sin(ang), 0.5 + h])
        tgt = np.array([rng.uniform(-0.5, 0.5), rng.uniform(-0.6, 0.6), rng.uniform(0.2, 0.9)])
        d = tgt - rel
        dxy = np.linalg.norm(d[:2])
        v, g = 6.0, 9.81
        num = v ** 4 - g * (g * dxy ** 2 + 2 * d[2] * v ** 2)
        theta = np.arctan((v ** 2 - np.sqrt(max(num, 0.0))) / (g * dxy))
        vel = np.array([v * np.cos(theta) * d[0] / dxy, v * np.cos(theta) * d[1] / dxy, v * np.sin(theta)])
        out[i] = np.concatenate([rel, vel, rng.uniform(-0.3, 0.3, 3), [rng.uniform(0, 2 * np.pi)],
                                 [float(rng.integers(150, 260))], [float(rng.integers(120, 260))]])
    return out


def record(name, cfg, seed):
    scene = Scene(cfg)
    rng = np.random.default_rng(seed)
    lo, hi = np.array(scene.pos_lo), np.array(scene.pos_hi)
    vmax, amax = np.array(scene.vel_max), np.array(scene.acc_max)
    q = rng.uniform(0.6 * lo, 0.6 * hi, (N, 7))
    q[:, 1] = rng.uniform(-0.9, 0.9, N)  # keep the arm off the table at the start
    q[:, 3] = rng.uniform(-1.2, 1.2, N)
    v = rng.uniform(-0.3, 0.3, (N, 7)) * vmax
    a = rng.uniform(-0.2, 0.2, (N, 7)) * amax
    ob = np.zeros((N, 16))
    is_ball = scene.struct.n_obstacles and scene.struct.obst_kind[0] == 2
    balls = None
    if is_ball:
        b = synth_ball(rng, N)
        ob[:, 2:12], ob[:, 14:16], ob[:, 13] = b[:, :10], b[:, 10:12], 1.0
        n0 = rng.integers(0, 3, N) * 24
        ob[:, 0], ob[:, 12] = n0, n0 * (0.1 / 24)
        balls = np.stack([synth_ball(rng, N) for _ in range(STEPS)])
    else:
        ob[:, 0] = rng.integers(0, scene.struct.planet_steps, N)
    env = oracle.OracleEnvs(scene, N)
    env.set_state(q, v, a, ob)
    actions = rng.uniform(-1, 1, (STEPS, N, 7)).astype(np.float32)
    actions[::5] = np.sign(actions[::5])  # saturated actions exercise the limit logic
    out = dict(q=q, v=v, a=a, obst=ob, actions=actions, obs0=env.obs.copy(), kin0=env.kin.copy())
    if balls is not None:
        out["balls"] = balls
    keys = ["kin", "obst", "obs", "reward", "done", "term", "info"]
    rec = {k: [] for k in keys}
    for s in range(STEPS):
        env.step(actions[s], balls[s] if balls is not None else None)
        for k, arr in zip(keys, [env.kin, env.obst, env.obs, env.reward, env.done, env.term, env.info]):
            rec[k].append(arr.copy())
    for k in keys:
        out["out_" + k] = np.stack(rec[k])
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    done = out["out_done"]
    print(name, "->", path, os.path.getsize(path), "bytes; first-done steps:",
          np.bincount(np.argmax(done > 0, axis=0), minlength=STEPS).tolist())


def record_human(name, cfg, seed, n=20, steps=20):
    """Human scene (BASELINE.json configs[2]): the nested env starts near its rest pose, the human's actions are inputs
    (the stochastic policy head is tested on its own), a reached target point is replaced by the recorded next one."""
    scene = Scene(cfg)
    rng = np.random.default_rng(seed)
    lo, hi = np.array(scene.pos_lo), np.array(scene.pos_hi)
    q = rng.uniform(0.5 * lo, 0.5 * hi, (n, 7))
    q[:, 1] = rng.uniform(-0.6, 0.3, n)
    q[:, 3] = rng.uniform(-1.2, 1.2, n)
    v = rng.uniform(-0.2, 0.2, (n, 7)) * np.array(scene.vel_max)
    a = np.zeros((n, 7))
    hlo, hhi = np.array(scene.human_pos_lo), np.array(scene.human_pos_hi)
    hq = np.zeros((n, 8))
    for e in range(n):                                         # random poses the braking-trajectory check accepts
        while True:
            cand = rng.uniform(0.8 * hlo + 0.2 * hhi, 0.2 * hlo + 0.8 * hhi)
            if not oracle.human_pose_collides(scene, cand):
                hq[e] = cand
                break
    hv = rng.uniform(-0.15, 0.15, (n, 8)) * np.array(scene.human_vel_max)
    ha = np.zeros((n, 8))
    arm = rng.integers(0, 2, n)
    box = np.array([[0.0, 0.6], [-0.8, 0.8], [0.075, 0.75]])   # target point box of the nested env (human params.json)
    ft = rng.uniform(box[:, 0], box[:, 1], (n, 3))
    env = oracle.OracleEnvs(scene, n)
    env.set_state(q, v, a, np.zeros((n, 16)))
    env.set_human_state(hq, hv, ha, ft, arm)
    actions = rng.uniform(-1, 1, (steps, n, 7)).astype(np.float32)
    hactions = rng.uniform(-1, 1, (steps, n, 8)).astype(np.float32)
    hactions[::4] = np.sign(hactions[::4])
    next_targets = rng.uniform(box[:, 0], box[:, 1], (steps, n, 3))
    out = dict(q=q, v=v, a=a, hq=hq, hv=hv, ha=ha, arm=arm, first_target=ft, actions=actions, hactions=hactions,
               next_targets=next_targets, obs0=env.obs.copy(), kin0=env.kin.copy(), hkin0=env.hkin.copy(),
               hobs0=env.hobs.copy(), hstate0=env.hstate.copy(), hbrake0=env.hbrake.copy())
    keys = ["kin", "hkin", "hstate", "hbrake", "obs", "hobs", "reward", "done", "term", "info", "hinfo"]
    rec = {k: [] for k in keys}
    for s in range(steps):
        env.step_human(actions[s], hactions[s], next_targets[s])
        for k, arr in zip(keys, [env.kin, env.hkin, env.hstate, env.hbrake, env.obs, env.hobs, env.reward, env.done,
                                 env.term, env.info, env.hinfo]):
            rec[k].append(arr.copy())
    for k in keys:
        out["out_" + k] = np.stack(rec[k])
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    done = out["out_done"]
    print(name, "->", path, os.path.getsize(path), "bytes; first-done steps:",
          np.bincount(np.argmax(done > 0, axis=0), minlength=steps).tolist(), "never:", int((done.max(0) == 0).sum()),
          "braked steps:", int((out["out_hinfo"][:, :, 0] != 0).sum()))


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    record("space", space_backup_config(), 11)
    record("ball", ball_backup_config(), 12)
    record("space_bm", space_backup_config(ball_machine_mode=True), 13)
    record("ball_bm", ball_backup_config(ball_machine_mode=True), 14)
    record_human("human", human_backup_config(), 15)
