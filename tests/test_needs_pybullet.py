"""Pins the oracle to the REAL reference.  Runs only where the reference's native dependencies
import (`pybullet`, `klimits`, `gym`): they are absent from the build container and from the GPU box and cannot be
installed there (no network), so this suite is skipped in both -- it exists so that any box that has them can decide
the "parity unpinned" question in one command:

    PYTHONPATH=/path/to/safeMotionsRisk python -m pytest tests/test_needs_pybullet.py -m needs_pybullet

Protocol: never compare RNG-driven resets.  The real env is reset, its start state (q, v, a, planet index /
ball launch) is read back and injected into the oracle, a fixed action sequence is applied to both, and every step's
(q, v, a), distances, collision flags, reward and observation are diffed at the north-star tolerances
(BASELINE.json: trajectories 1e-5 rad, distances 1e-4 m, rewards 1e-4 relative, flags exact except contacts within
1e-5 m of a threshold).  The traces recorded from the real env are written to tests/golden/reference_<scene>.npz so
that they can be committed as reference-held golden vectors.
"""
import importlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.needs_pybullet


def _reference_available():
    for mod in ("pybullet", "klimits", "gym"):
        try:
            importlib.import_module(mod)
        except Exception:
            return False
    try:
        importlib.import_module("safemotions.envs.safe_motions_env")
    except Exception:
        return False
    return True


needs_ref = pytest.mark.skipif(not _reference_available(),
                               reason="pybullet / klimits / gym / safemotions not importable (expected here)")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _reference_env(cfg):
    from safemotions.envs.safe_motions_env import SafeMotionsEnvCollisionAvoidance
    kwargs = {k: v for k, v in cfg.items() if k != "contact_check_stride"}
    kwargs.setdefault("experiment_name", "pin")
    kwargs["seed"] = 0
    return SafeMotionsEnvCollisionAvoidance(**kwargs)


def _read_start_state(env, scene):
    """(q, v, a, obstacle record) of the real env right after reset()."""
    tm = env._trajectory_manager
    q = np.array(tm.get_trajectory_start_position(), dtype=np.float64) if hasattr(tm, "get_trajectory_start_position") \
        else np.array(env._start_position, dtype=np.float64)
    v = np.array(env._start_velocity, dtype=np.float64)
    a = np.array(env._start_acceleration, dtype=np.float64)
    ob = np.zeros(16)
    wrapper = env._robot_scene.obstacle_wrapper
    if wrapper.planet_list:
        ob[0] = wrapper.planet_list[0].current_time_step_index
    elif env._robot_scene.use_moving_objects:
        ball = [b for lst in wrapper._moving_object_list for b in lst][-1]
        ob[2:5], ob[5:8] = ball._base_pos, ball._initial_speed_vector
        ob[8:11], ob[11] = ball._initial_orn_euler, ball._angular_velocity_euler_y if hasattr(
            ball, "_angular_velocity_euler_y") else 0.0
        ob[12], ob[13] = ball._t, 1.0
        ob[14], ob[15] = ball.max_time_update_step_counter, ball.obstacle_hit_time
    return q, v, a, ob


@needs_ref
@pytest.mark.parametrize("name", ["space", "ball"])
def test_oracle_matches_the_real_env(name):
    from oracle import oracle
    from safemotionsrisk_b200 import abi, ball_backup_config, space_backup_config
    from safemotionsrisk_b200.scene import Scene
    cfg = dict(space_backup_config() if name == "space" else ball_backup_config())
    ref = _reference_env(cfg)
    scene = Scene(space_backup_config() if name == "space" else ball_backup_config())
    rng = np.random.default_rng(0)
    traces = []
    for episode in range(8):
        ref.reset()
        q, v, a, ob = _read_start_state(ref, scene)
        orc = oracle.OracleEnvs(scene, 1)
        orc.set_state(q[None], v[None], a[None], ob[None])
        for step in range(scene.struct.episode_steps):
            u = rng.uniform(-1, 1, 7).astype(np.float32)
            r_obs, r_rew, r_done, r_info = ref.step(u.astype(np.float64))
            o_obs, o_rew, o_done, _, o_info = orc.step(u[None], None)
            traces.append(dict(q=np.array(ref._start_position), v=np.array(ref._start_velocity),
                               a=np.array(ref._start_acceleration), obs=np.array(r_obs), reward=r_rew, done=r_done))
            assert np.abs(orc.kin[0, 0:7] - np.array(ref._start_position)).max() < 1e-5, "joint positions"
            assert np.abs(orc.kin[0, 8:15] - np.array(ref._start_velocity)).max() < 1e-5, "joint velocities"
            assert np.abs(orc.kin[0, 16:23] - np.array(ref._start_acceleration)).max() < 1e-4, "safe-action clipping"
            assert np.abs(o_obs[0] - np.asarray(r_obs)).max() < 1e-4, "observation"
            edge = min(abs(o_info[0, c] - t) for c in range(3) for t in (1e-3,)) < 1e-5
            if not edge:
                assert bool(o_done[0]) == bool(r_done), "termination"
                assert abs(o_rew[0] - r_rew) <= 1e-4 * max(1.0, abs(r_rew)), "reward"
            if r_done:
                break
    np.savez_compressed(os.path.join(GOLDEN, "reference_{}.npz".format(name)),
                        **{"{}_{}".format(k, i): np.asarray(t[k]) for i, t in enumerate(traces) for k in t})


@needs_ref
def test_safe_range_matches_klimits():
    """klimits.PosVelJerkLimitation with the constructor arguments of actions.py:97-106 against the oracle's range on
    random states: the north star wants the clipping bit-exact; this reports the largest difference."""
    from klimits import PosVelJerkLimitation
    from oracle import oracle
    from safemotionsrisk_b200 import space_backup_config
    from safemotionsrisk_b200.scene import Scene
    scene = Scene(space_backup_config())
    lim = PosVelJerkLimitation(time_step=0.1, pos_limits=[[lo, hi] for lo, hi in zip(scene.pos_lo, scene.pos_hi)],
                               vel_limits=[[-v, v] for v in scene.vel_max], acc_limits=[[-a, a] for a in scene.acc_max],
                               jerk_limits=[[-j, j] for j in scene.jerk_max],
                               acceleration_after_max_vel_limit_factor=0.01, set_velocity_after_max_pos_to_zero=True,
                               limit_velocity=True, limit_position=True, normalize_acc_range=False)
    rng = np.random.default_rng(1)
    worst = 0.0
    for _ in range(2000):
        q = rng.uniform(scene.pos_lo, scene.pos_hi)
        v = rng.uniform(-1, 1, 7) * scene.vel_max * 0.8
        a = rng.uniform(-1, 1, 7) * scene.acc_max * 0.5
        ref_range, _ = lim.calculate_valid_acceleration_range(q, v, a)
        lo, hi, code = oracle.safe_range(scene, q, v, a)
        if not np.any(code):
            worst = max(worst, np.abs(np.asarray(ref_range)[:, 0] - lo).max(), np.abs(np.asarray(ref_range)[:, 1] - hi).max())
    assert worst < 1e-6, "largest range difference {:.3e} rad/s^2".format(worst)
