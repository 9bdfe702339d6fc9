"""Scene builder against the constants of the reference."""
import numpy as np
import pytest

from safemotionsrisk_b200 import EnvConfig, ball_backup_config, space_backup_config
from safemotionsrisk_b200.scene import Scene, planet_tables


def test_limits_match_reference_constants(space_scene):
    # robot_scene_base.py:20-22, :347-371, :441-455 at factor 1.0, dt = 0.1
    urdf_upper = np.array([2.9670597283903604, 2.0943951023931953, 2.9670597283903604, 2.0943951023931953,
                           2.9670597283903604, 2.0943951023931953, 3.0543261909900763])
    assert np.allclose(space_scene.pos_hi, urdf_upper - 0.035)
    assert np.allclose(space_scene.pos_lo, -(urdf_upper - 0.035))
    assert np.allclose(space_scene.vel_max, [1.710422666954443] * 2 + [1.7453292519943295, 2.2689280275926285,
                                                                        2.443460952792061, 3.141592653589793,
                                                                        3.141592653589793])
    assert np.allclose(space_scene.acc_max, [15, 7.5, 10, 12.5, 15, 20, 20])
    assert np.allclose(space_scene.jerk_max, [300, 150, 200, 250, 300, 400, 400])


def test_step_structure(space_scene, ball_scene):
    for s in (space_scene, ball_scene):
        assert s.struct.substeps == 24 and s.struct.episode_steps == 20 and s.struct.n_joints == 7
    assert space_scene.obs_size == 23 and ball_scene.obs_size == 27          # risk_config.json observation_size
    assert abs(space_scene.struct.static_cap - 0.102) < 1e-12               # ctlp.py:354-366
    assert space_scene.struct.moving_query == 0.6


def test_pair_sets(space_scene, space_bm_scene, ball_scene):
    # table x links 2-7; planets x links 3-7; ball x links 2-7; ball machine x 7 lower links
    assert space_scene.struct.n_static_pairs == 6 and space_scene.struct.n_self_pairs == 0
    assert space_scene.struct.n_mov_reward == 5 and space_scene.struct.obst_shape_cnt[0] == 24
    assert space_scene.struct.obst_shape_cnt[1] == 1
    assert ball_scene.struct.n_mov_reward == 6 and ball_scene.struct.obst_shape_cnt[0] == 1
    assert space_bm_scene.struct.n_static_pairs == 8 and space_bm_scene.struct.n_self_pairs == 14
    assert space_bm_scene.struct.n_mov_reward == 7


def test_planet_tables_appendix_c():
    upd = 0.1 / 24
    pos1, quat1, loc1, len1 = planet_tables([-0.1, 0, 0.8], [0.65, 0.8], [0.35, 0, 0], 5.0, upd, [0, 0, -np.pi / 2], 1)
    pos2, _, _, len2 = planet_tables([-0.1, 0, 0.8], [0.75, 0.8], [-0.35, 0, 0], 5.0, upd, [0, 0, 0], 4)
    assert len(pos1) == 1200 and len(pos2) == 1200
    assert abs(len1 - 4.5675) < 1e-3 and abs(len2 - 4.8707) < 1e-3
    hop = np.linalg.norm(np.diff(pos1, axis=0), axis=1).max()
    assert 4.5e-3 < hop < 5.5e-3                                             # ~1.2 m/s at 240 Hz
    assert np.allclose(np.linalg.norm(quat1, axis=1), 1.0)
    # the orbit lies in the tilted plane through the centre
    n = np.array([0, -np.sin(0.35), np.cos(0.35)])
    assert np.abs((pos1 - [-0.1, 0, 0.8]) @ n).max() < 1e-12
    assert np.allclose(loc1[0], [0.65, 0.0])


def test_planet_shift(space_scene):
    assert space_scene.struct.planet_shift == -480                           # ctlp.py:4435-4437


def test_ball_observation_ranges(ball_scene):
    sc = ball_scene.struct
    assert np.allclose(list(sc.ball_obs_pos_min), [-2.5, -2.5, -0.3])
    assert np.allclose(list(sc.ball_obs_pos_max), [2.5, 2.5, 2.8349], atol=1e-4)
    assert np.allclose(list(sc.ball_obs_vel_max), [6, 6, 7.4077], atol=1e-4)


def test_unknown_keys_are_swallowed_and_unsupported_raise():
    cfg = EnvConfig(floating_robot_base=False, num_virtual_motors=0, action_max_reward=1.0)   # Appendix A, Q13
    assert "floating_robot_base" in cfg.ignored
    with pytest.raises(NotImplementedError):
        EnvConfig(check_braking_trajectory_collisions=True)
    with pytest.raises(NotImplementedError):
        EnvConfig(robot_scene=2)                                              # not defined by the reference either
    with pytest.raises(NotImplementedError):
        Scene(space_backup_config(human_network_checkpoint="human_network/checkpoint/checkpoint"))
    with pytest.raises(ValueError):
        Scene(space_backup_config(closest_point_safety_distance=0.2))         # base.py-style config error


def test_shipped_params_json_keys_load():
    # the env_config of trained_networks/backup_networks/space/params.json (keys copied from the survey)
    cfg = space_backup_config(ball_machine_mode=True, target_link_offset=[0, 0, 0], ray_version="1.4.1",
                              klimits_version="1.1.3", floating_robot_base=False, episodes_per_simulation_reset=4000)
    s = Scene(cfg)
    assert s.target_link == "ball_machine"


def test_contact_thresholds_are_millimetres(space_scene, ball_scene):
    for s in (space_scene, ball_scene):
        th = np.array([[s.struct.contact_thresh[o][i] for i in range(s.struct.n_mov_contact)]
                       for o in range(s.struct.n_obstacles)])
        assert (th > 5e-4).all() and (th < 1e-2).all()


@pytest.mark.parametrize("key,value", [
    ("acceleration_after_max_vel_limit_factor", 0.05), ("set_velocity_after_max_pos_to_zero", False),
    ("punish_adaptation", True), ("punish_end_min_distance", True), ("punish_end_max_torque", True),
    ("punish_braking_trajectory_min_distance", True), ("punish_braking_trajectory_max_torque", True),
    ("obstacle_use_computed_actual_values", True), ("check_braking_trajectory_torque_limits", True),
    ("risk_state_config", 2), ("risk_store_ground_truth", True), ("human_network_use_full_observation", True),
    ("terminate_on_robot_stop", True), ("max_resampling_attempts", 3), ("action_preprocessing_function", "clip"),
])
def test_keys_that_change_unimplemented_behaviour_raise(key, value):
    """A params.json that sets one of these keys away from the reference default would silently get different rewards
    or limits: every one of them raises instead."""
    with pytest.raises(NotImplementedError):
        EnvConfig(**{key: value})
    EnvConfig(**{key: EnvConfig()[key]})   # the default value itself is accepted


def test_accepted_keys_have_an_effect_or_are_no_ops():
    assert Scene(space_backup_config(normalize_reward_to_frequency=True, trajectory_time_step=0.05)
                 ).struct.reward_scale == pytest.approx(0.5)
    assert Scene(space_backup_config()).struct.reward_scale == 1.0
    EnvConfig(action_preprocessing_function="tanh")     # computed and discarded by the reference
    EnvConfig(no_self_collision=True)                    # a Bullet loading flag of the dynamics, which are not modelled
    with pytest.raises(ValueError):
        EnvConfig(risk_config_dir="risk_networks/state_action/space")   # needs risk_threshold
