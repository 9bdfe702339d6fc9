"""Flag surface of train.py / evaluate.py (train.py:168-309, :347-513): README commands map onto the env_config."""
import pytest

from safemotionsrisk_b200 import space_backup_config
from safemotionsrisk_b200.cli import add_env_arguments, env_config_from_args, env_config_from_command

# README.md:74 (backup policy, Space)
SPACE = ("python safemotions/train.py --logdir=specify_path_to_store_checkpoints --name=backup_space --robot_scene=0 "
         "--planet_mode --planet_one_center=\"[-0.1,0.0,0.8]\" --planet_one_euler_angles=\"[0.35,0,0]\" "
         "--planet_one_period=5.0 --planet_one_radius_xy=\"[0.65,0.8]\" --planet_two_center=\"[-0.1,0,0.8]\" "
         "--planet_two_euler_angles=\"[-0.35,0,0]\" --planet_two_radius_xy=\"[0.75,0.8]\" --planet_two_time_shift=-2.0 "
         "--obs_planet_size_per_planet=2 --obstacle_scene=5 --terminate_on_collision_with_moving_obstacle "
         "--terminate_on_collision_with_static_obstacle --terminate_on_self_collision --collision_avoidance_mode "
         "--collision_avoidance_kinematic_state_sampling_mode --collision_avoidance_kinematic_state_sampling_probability=0.7 "
         "--collision_avoidance_stay_in_state_probability=0.3 --collision_avoidance_moving_obstacles_max_reward=3.0 "
         "--collision_avoidance_moving_obstacles_max_reward_distance=0.6 --collision_avoidance_self_collision_max_reward=1.0 "
         "--collision_avoidance_static_obstacles_max_reward=1.0 --collision_avoidance_low_acceleration_max_reward=0.0 "
         "--collision_avoidance_low_velocity_max_reward=0.0 --collision_avoidance_episode_termination_bonus=15.0 "
         "--collision_avoidance_episode_early_termination_punishment=-15.0 --closest_point_safety_distance=0.01 "
         "--use_controller_target_velocities --starting_point_cartesian_range_scene=1 --punish_action "
         "--action_punishment_min_threshold=0.95 --action_max_punishment=0.4 --collision_check_time=0.033 "
         "--trajectory_duration=2.0 --solver_iterations=50 --num_workers=12 --num_gpus=1 --time=500").replace('"', "")


def test_readme_command_gives_the_space_backup_config():
    cfg = env_config_from_command(SPACE)
    ref = space_backup_config()
    for key in ref:
        if key in ("collision_avoidance_low_acceleration_threshold", "collision_avoidance_low_velocity_threshold"):
            continue   # thresholds only matter with a non-zero weight; the shipped params.json differs from the flag default
        assert cfg[key] == ref[key], key
    assert cfg.experiment_name == "backup_space"


def test_flags_cover_the_env_keywords_and_unknown_ones_are_ignored():
    parser = add_env_arguments()
    args, rest = parser.parse_known_args(["--ball_machine_mode", "--risk_threshold=0.065",
                                          "--risk_config_dir=risk_networks/state_action/space",
                                          "--risk_state_config=RISK_CHECK_CURRENT_STATE", "--unknown_flag=3"])
    cfg = env_config_from_args(args)
    assert cfg.ball_machine_mode and cfg.risk_threshold == 0.065 and cfg.risk_state_config == 0
    assert rest == ["--unknown_flag=3"]
    with pytest.raises(NotImplementedError):   # a flag that selects an unimplemented part fails loudly
        env_config_from_command("--check_braking_trajectory_torque_limits")
