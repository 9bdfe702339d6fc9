"""Oracle env step: golden regression vectors, reward algebra, termination priority, obstacle bookkeeping."""
import os

import numpy as np
import pytest

from oracle import oracle
from safemotionsrisk_b200 import abi, ball_backup_config, space_backup_config
from safemotionsrisk_b200.scene import Scene

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CONFIGS = {"space": lambda: space_backup_config(), "ball": lambda: ball_backup_config(),
           "space_bm": lambda: space_backup_config(ball_machine_mode=True),
           "ball_bm": lambda: ball_backup_config(ball_machine_mode=True)}


@pytest.mark.parametrize("name", ["space", "ball", "space_bm", "ball_bm"])
def test_oracle_reproduces_golden_vectors(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    scene = Scene(CONFIGS[name]())
    n = g["q"].shape[0]
    steps = 6 if name.startswith("space") else 20   # the brute-force Space oracle is slow; 6 steps pin it
    env = oracle.OracleEnvs(scene, n)
    env.set_state(g["q"], g["v"], g["a"], g["obst"])
    assert np.array_equal(env.obs, g["obs0"]) and np.array_equal(env.kin, g["kin0"])
    for s in range(steps):
        env.step(g["actions"][s], g["balls"][s] if "balls" in g else None)
        assert np.array_equal(env.kin, g["out_kin"][s])
        assert np.array_equal(env.obst, g["out_obst"][s])
        assert np.array_equal(env.obs, g["out_obs"][s])
        assert np.array_equal(env.reward, g["out_reward"][s])
        assert np.array_equal(env.done, g["out_done"][s]) and np.array_equal(env.term, g["out_term"][s])


def test_reward_algebra_backup_config(space_scene):
    """rewards.py:432-502 with the README weights: 0.4 (1 - pun) + r_self + r_static + 3 r_moving (+15 | -15)."""
    n = 8
    rng = np.random.default_rng(4)
    env = oracle.OracleEnvs(space_scene, n)
    q = np.tile([0.0, 0.4, 0.0, -0.8, 0.0, 0.5, 0.0], (n, 1)) + rng.uniform(-0.1, 0.1, (n, 7))
    ob = np.zeros((n, 16))
    ob[:, 0] = rng.integers(0, 1200, n)
    env.set_state(q, np.zeros((n, 7)), np.zeros((n, 7)), ob)
    u = rng.uniform(-1, 1, (n, 7)).astype(np.float32)
    u[0] = 0.99                                                  # saturated action: punished
    _, rew, done, term, info = env.step(u)
    I = abi.INFO
    pun = np.clip((np.abs(u).max(1) - 0.95) / 0.05, 0, 1) ** 2   # rewards.py:18-21, :164-169
    assert np.allclose(info[:, I["action_punishment"]], pun, atol=1e-6)
    rel = lambda d, dmax: np.minimum(1.0, d / dmax) ** 2
    expect = 0.4 * (1 - pun) + rel(info[:, I["d_self"]], 0.05) + rel(info[:, I["d_static"]], 0.1) + \
        3 * rel(info[:, I["d_moving"]], 0.6)
    coll = (info[:, I["coll_static"]] + info[:, I["coll_self"]] + info[:, I["coll_moving"]]) > 0
    expect = expect + np.where(coll, -15.0, 0.0)
    assert np.allclose(rew, expect, atol=1e-5)
    assert (rew <= 5.4 + 1e-6).all()                              # Appendix C: max per-step reward
    assert np.array_equal(done > 0, coll)


def test_episode_length_and_termination_bonus():
    scene = Scene(space_backup_config(contact_check_stride=0, terminate_on_collision_with_moving_obstacle=False,
                                      terminate_on_collision_with_static_obstacle=False))
    env = oracle.OracleEnvs(scene, 2)
    ob = np.zeros((2, 16))
    env.set_state(np.zeros((2, 7)), np.zeros((2, 7)), np.zeros((2, 7)), ob)
    rewards = []
    for s in range(20):
        _, rew, done, term, _ = env.step(np.zeros((2, 7), dtype=np.float32))
        rewards.append(rew.copy())
        assert (done > 0).all() == (s == 19)
    assert (term == abi.TERMINATION_TRAJECTORY_LENGTH).all()      # trajectory_manager.py:187-192
    assert np.allclose(rewards[-1] - rewards[-2], 15.0, atol=0.3)  # bonus only at the last step


def test_termination_priority_static_before_moving(space_scene):
    # arm folded into the table AND a latched planet contact: static (4) wins over moving (5) (base.py:1775-1799)
    env = oracle.OracleEnvs(space_scene, 1)
    ob = np.zeros((1, 16))
    ob[0, 1] = 1.0
    env.set_state(np.array([[1.3, -2.0, 1.5, 0.1, 2.5, -1.8, 2.1]]), np.zeros((1, 7)), np.zeros((1, 7)), ob)
    _, rew, done, term, info = env.step(np.zeros((1, 7), dtype=np.float32))
    assert done[0] and term[0] == abi.TERMINATION_COLLISION_WITH_STATIC_OBSTACLE
    assert info[0, abi.INFO["coll_moving"]] == 1 and info[0, abi.INFO["d_moving"]] == 0
    assert rew[0] < -10


def test_planet_index_advances_24_per_step(space_scene):
    env = oracle.OracleEnvs(space_scene, 1)
    ob = np.zeros((1, 16))
    ob[0, 0] = 1190
    env.set_state(np.zeros((1, 7)), np.zeros((1, 7)), np.zeros((1, 7)), ob)
    env.step(np.zeros((1, 7), dtype=np.float32))
    assert env.obst[0, 0] == (1190 + 24) % 1200                    # 50 env steps per orbit (Appendix C)


def test_planet_observation_is_local_orbit_xy(space_scene):
    ob = np.zeros(16)
    kin = np.zeros(32)
    obs = oracle.observation(space_scene, kin, ob)
    assert obs.shape == (23,)
    assert abs(obs[21] - 0.65 / (1.05 * 0.65)) < 1e-6 and abs(obs[22]) < 1e-6   # observations.py:280-288
    kin[0] = space_scene.pos_hi[0]
    kin[8] = -space_scene.vel_max[0]
    obs = oracle.observation(space_scene, kin, ob)
    assert obs[0] == 1.0 and obs[7] == -1.0


def test_ball_bookkeeping(ball_scene):
    env = oracle.OracleEnvs(ball_scene, 3)
    ob = np.zeros((3, 16))
    ob[:, 2:5] = [2.4, 0.3, 1.0]
    ob[:, 5:8] = [-5.5, -0.6, 1.5]
    ob[:, 13] = 1
    ob[:, 14] = [200, 10, 200]      # env 1 leaves the box during the step (nmax)
    ob[:, 15] = [150, 150, 30]      # env 2 hits the table/ground at counter 30
    ob[:, 0] = [0, 0, 24]
    ob[:, 12] = ob[:, 0] * (0.1 / 24)
    env.set_state(np.zeros((3, 7)), np.zeros((3, 7)), np.zeros((3, 7)), ob)
    env.step(np.zeros((3, 7), dtype=np.float32))
    assert env.obst[0, 0] == 24 and env.obst[0, 13] == 1 and abs(env.obst[0, 12] - 0.1) < 1e-12
    assert env.obst[1, 0] == 11 and env.obst[1, 13] == 0           # counter > nmax at sub-step 11 (ctlp.py:4196)
    assert env.obst[2, 0] == 30 and env.obst[2, 13] == 0           # counter >= hit time (ctlp.py:4207-4211)
    assert env.info[1, abi.INFO["d_moving"]] == np.float32(0.602)   # inactive ball: cap value (ctlp.py:3259-3261)
    # a replacement launch is taken over at observation time
    nb = np.tile(np.array([2.0, 1.0, 0.8, -5.0, -2.0, 2.0, 0.1, 0.2, 0.3, 1.0, 180.0, 160.0]), (3, 1))
    env.step(np.zeros((3, 7), dtype=np.float32), nb)
    assert env.obst[1, 13] == 1 and env.obst[1, 0] == 0 and np.allclose(env.obst[1, 2:5], [2.0, 1.0, 0.8])
    assert abs(env.obs[1, 21] - (-1 + 2 * (2.0 + 2.5) / 5.0)) < 1e-6   # ball position observation (ctlp.py:2364-2371)


def test_philox_known_answer():
    # Random123 known-answer test for philox4x32-10
    assert oracle.philox(0, 0, 0, 0, 0, 0) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff) == \
        [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]


def test_target_point_reward_and_replacement():
    """TargetPointReachingReward on the oracle (rewards.py:303-396; ctlp.py:2309-2350): progress reward
    (last - current) / (ts * initial), reached reward last / (ts * initial) + bonus, replacement by the next target."""
    from safemotionsrisk_b200.config import space_task_config
    sc = Scene(space_task_config(punish_action=False))
    env = oracle.OracleEnvs(sc, 2)
    q = np.tile(np.array([0.3, 0.5, -0.2, -1.0, 0.1, 0.8, 0.0]), (2, 1))
    z = np.zeros_like(q)
    ob = np.zeros((2, 16))
    ob[:, 0] = 300           # planets far from this pose
    link = oracle.target_link_point(sc, q[0])
    ft = np.stack([link + [0.2, 0.0, 0.0], link + [0.03, 0.0, 0.0]])   # env 1 starts inside the radius (6.5 cm)
    env.set_state(q, z, z, ob, ft)
    assert env.obs.shape[1] == 29
    assert np.allclose(env.tp[:, 3], [0.2, 0.03]) and np.allclose(env.tp[:, 4], [0.2, 0.03])
    nxt = np.tile([0.3, 0.3, 0.5], (2, 1))
    obs, rew, done, term, info = env.step(np.zeros((2, 7), dtype=np.float32), None, nxt)
    # env 1: reached in the first sub-step -> 0.03 / (0.1 * 0.03) + 5
    assert abs(rew[1] - 15.0) < 1e-4 and env.tp[1, 6] == 1 and np.allclose(env.tp[1, 0:3], nxt[1])
    # env 0: progress reward from the distances before / after the step
    d_after = np.linalg.norm(ft[0] - env.tp[0, 7:10])
    assert abs(rew[0] - (0.2 - d_after) / (0.1 * 0.2)) < 1e-4
    assert abs(env.tp[0, 3] - d_after) < 1e-12 and abs(env.tp[0, 4] - 0.2) < 1e-12   # last updated, initial kept
    # observation: target position and relative position, normalised (observations.py:326-340)
    lo, hi = np.array(sc.struct.tp_box_min), np.array(sc.struct.tp_box_max)
    assert np.allclose(obs[0, 21:24], np.clip(-1 + 2 * (ft[0] - lo) / (hi - lo), -1, 1), atol=1e-6)
    rlo, rhi = np.array(sc.struct.tp_rel_min), np.array(sc.struct.tp_rel_max)
    assert np.allclose(obs[0, 24:27], -1 + 2 * ((ft[0] - env.tp[0, 7:10]) - rlo) / (rhi - rlo), atol=1e-6)
