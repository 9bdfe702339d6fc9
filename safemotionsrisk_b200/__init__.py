"""B200-native implementation of the per-RL-step environment hot path of translearn/safeMotionsRisk."""
from .config import (EnvConfig, ball_backup_config, human_backup_config, space_backup_config,  # noqa: F401
                     space_task_config)

__all__ = ["EnvConfig", "space_backup_config", "ball_backup_config", "space_task_config", "human_backup_config", "SafeMotionsVecEnv", "make_env"]


def __getattr__(name):  # torch is imported only when the env class is requested
    if name in ("SafeMotionsVecEnv", "make_env"):
        from . import vec_env
        return getattr(vec_env, name)
    raise AttributeError(name)
