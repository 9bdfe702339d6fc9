#!/usr/bin/env python
"""GJK work counters per env-step from the counting build of the phase kernels.
Usage (GPU box): python tools/phase_profile.py [scene ...]  -- kernel time shares come from the ncu launch list."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import scene_config
from safemotionsrisk_b200.vec_env import SafeMotionsVecEnv
for scene in (sys.argv[1:] or ["ball", "space"]):
    env = SafeMotionsVecEnv(num_envs=65536, config=scene_config(scene), seed=1)
    env.reset()
    for _ in range(25): env.step_random()
    env.enable_counters(True); env.counters(reset=True)
    for _ in range(5): env.step_random()
    c = env.counters(); n = c["env_steps"]
    print(scene, "per env-step: gjk pairs %.2f (distance items %.2f, contact items %.2f) iters %.2f dots %.0f" % (
        c["gjk_calls"]/n, c["distance_items"]/n, c["contact_items"]/n, c["gjk_iters"]/n, c["support_dots"]/n))
    print("   heavy joints %.2f (solves %.2f) per env-step, envs passed to the fine contact planning %.3f" % (
        c["heavy_joints"]/n, c["heavy_solves"]/n, c["contact_envs"]/n))
    print("   gjk pairs by iterations (<=4, 8, 12, 16, 24, more):", [round(x / max(1, c["gjk_calls"]), 4) for x in c["gjk_iteration_histogram"]])
    env.enable_counters(False); env.kernel_timing(True)
    for _ in range(20): env.step_random()
    t, k = env.kernel_times()
    print("   kernel us/step:", {a: round(1e3*b, 1) for a, b in t.items()}, "sum %.1f" % (1e3*sum(t.values())))
    env.close()
