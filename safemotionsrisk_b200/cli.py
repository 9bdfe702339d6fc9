"""The env part of the reference's command line: ``train.py`` / ``evaluate.py`` flags -> ``env_config``.

The reference builds its ``env_config`` dict from argparse flags that carry the names of the env's constructor keywords
(train.py:168-309 ``_make_env_config``, flags train.py:347-513; evaluate.py:452-639 patches a checkpoint's config with the
same names).  This module re-creates that flag surface from the key table of ``config.py`` so that the README's training
commands (README.md:74-95, :223-235) can be pasted unchanged:

    parser = argparse.ArgumentParser(); add_env_arguments(parser)
    args, _ = parser.parse_known_args("--planet_mode --obstacle_scene=5 ...".split())
    env = SafeMotionsVecEnv(num_envs=65536, config=env_config_from_args(args))

Only the env flags are mapped: the PPO / Ray flags of the scripts (--num_workers, --num_gpus, --batch_size_factor,
--logdir, --time, ...) are accepted and ignored, they configure the learner, which is out of scope.
"""
import argparse
import json

from .config import EnvConfig, _DEFAULTS

# train.py:33-37
RISK_STATE_CONFIG = {"RISK_CHECK_CURRENT_STATE": 0, "RISK_CHECK_NEXT_STATE_KINEMATIC_FORECASTING": 1,
                     "RISK_CHECK_NEXT_STATE_FULL_FORECASTING": 2, "RISK_CHECK_NEXT_STATE_SIMULATE_NEXT_STEP": 3,
                     "RISK_CHECK_NEXT_STATE_SIMULATE_NEXT_STEP_AND_BACKUP_TRAJECTORY": 4}
# flags of the scripts that do not reach the env (learner, logging): parsed so that pasted commands do not fail
_LEARNER_FLAGS = ["logdir", "checkpoint", "time", "iterations_per_checkpoint", "num_workers", "num_threads_per_worker",
                  "num_gpus", "batch_size_factor", "last_layer_activation", "log_std_range", "hidden_layer_activation",
                  "fcnet_hiddens", "episodes", "store_metrics", "use_gui"]
# env flags of train.py whose keywords the env swallows without this package knowing them (reward terms of the
# braking-trajectory machinery, Bullet options): kept in EnvConfig.ignored exactly like the reference's **kwargs
_EXTRA_ENV_FLAGS = dict(adaptation_max_punishment=1.0, end_min_distance_max_threshold=0.05, end_min_distance_max_punishment=1.0,
                        end_max_torque_min_threshold=0.9, end_max_torque_max_punishment=1.0,
                        braking_trajectory_min_distance_max_threshold=0.05, braking_trajectory_max_punishment=1.0,
                        braking_trajectory_max_torque_min_threshold=0.8, acc_limit_factor_braking=1.0,
                        jerk_limit_factor_braking=1.0, risk_ground_truth_episodes_per_file=None)


def _add(parser, key, default):
    flag = "--" + key
    if isinstance(default, bool):
        parser.add_argument(flag, action="store_true", default=default)
    elif isinstance(default, int):
        parser.add_argument(flag, type=int, default=default)
    elif isinstance(default, float):
        parser.add_argument(flag, type=float, default=default)
    elif isinstance(default, str):
        parser.add_argument(flag, type=str, default=default)
    else:   # None: numbers, lists and strings all arrive as JSON where possible (train.py uses type=json.loads for lists)
        def parse(text):
            try:
                return json.loads(text)
            except ValueError:
                return text
        parser.add_argument(flag, type=parse, default=default)


def add_env_arguments(parser=None):
    """Adds one flag per env keyword (train.py:347-513 names and defaults) and returns the parser."""
    parser = parser or argparse.ArgumentParser()
    parser.add_argument("--name", type=str, default="default")            # -> experiment_name (train.py:171)
    for key, default in list(_DEFAULTS.items()) + list(_EXTRA_ENV_FLAGS.items()):
        if key in ("experiment_name", "risk_state_config", "use_gui", "contact_check_stride"):
            continue
        _add(parser, key, default)
    parser.add_argument("--risk_state_config", default="RISK_CHECK_CURRENT_STATE", choices=list(RISK_STATE_CONFIG))
    for key in _LEARNER_FLAGS:
        parser.add_argument("--" + key, nargs="?", default=None)
    return parser


def env_config_from_args(args):
    """``_make_env_config`` of train.py:168-309: the parsed flags as an ``EnvConfig``."""
    values = vars(args)
    cfg = {key: values[key] for key in list(_DEFAULTS) + list(_EXTRA_ENV_FLAGS) if key in values}
    cfg["experiment_name"] = values.get("name", "default")
    cfg["risk_state_config"] = RISK_STATE_CONFIG[values.get("risk_state_config", "RISK_CHECK_CURRENT_STATE")]
    cfg["use_gui"] = bool(values.get("use_gui"))
    return EnvConfig(**cfg)


def env_config_from_command(command):
    """Convenience: the env_config of a pasted command line (e.g. the README's `python safemotions/train.py ...`)."""
    tokens = [t for t in command.split() if t.startswith("--")]
    args, _ = add_env_arguments().parse_known_args(tokens)
    return env_config_from_args(args)
