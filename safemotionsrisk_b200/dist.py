"""Multi-GPU plumbing: environments shard trivially, one process per GPU, no collective on the step path.

The reference runs one env per Ray rollout worker (train.py:598-608) and aggregates episode info on the driver
(CustomTrainCallbacks, train.py:40-120).  Here every rank owns a contiguous block of environments; the only collective
is the all-reduce of the episode-statistics vector once per training iteration (NCCL on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist

STAT_EPISODES, STAT_RETURN_SUM, STAT_LENGTH_SUM, STAT_REASON0 = 0, 1, 2, 3


def shard_range(total_envs, rank, world):
    """Contiguous block [start, stop) of the global env index space owned by `rank` (blocks differ by at most one)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world: {}/{}".format(rank, world))
    base, rem = divmod(int(total_envs), world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def rank_seed(seed, rank):
    """Seed of a shard: the bench uses 1000 * rank + seed."""
    return 1000 * int(rank) + int(seed)


def allreduce_episode_stats(stats):
    """Sums the per-shard statistics vector over all ranks (in place) and returns the custom_metrics the reference's
    callbacks report (train.py:59-117): episode count, mean return, mean length, termination-reason rates."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    s = stats.detach().cpu().double()
    n = max(float(s[STAT_EPISODES]), 1.0)
    out = {"episodes": float(s[STAT_EPISODES]), "episode_reward_mean": float(s[STAT_RETURN_SUM]) / n,
           "episode_len_mean": float(s[STAT_LENGTH_SUM]) / n}
    names = {1: "joint_limits", 2: "trajectory_length", 3: "self_collision", 4: "collision_with_static_obstacle",
             5: "collision_with_moving_obstacle"}
    for r, name in names.items():
        out["termination_reason_{}_rate".format(name)] = float(s[STAT_REASON0 + r]) / n
    return out
