"""Host-side multi-rank logic on CPU (gloo, world_size 2): shard ranges and the episode-statistics all-reduce."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from safemotionsrisk_b200.dist import allreduce_episode_stats, rank_seed, shard_range


def test_shard_ranges_cover_everything_once():
    for total, world in ((65536, 8), (10, 3), (7, 8), (131072, 4)):
        ranges = [shard_range(total, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == total
        for (a0, a1), (b0, b1) in zip(ranges[:-1], ranges[1:]):
            assert a1 == b0 and 0 <= (a1 - a0) - (b1 - b0) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)
    assert rank_seed(0, 3) == 3000


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    stats = torch.zeros(32, dtype=torch.float64)
    stats[0] = 10 * (rank + 1)           # episodes
    stats[1] = 50.0 * (rank + 1)         # return sum
    stats[2] = 200.0 * (rank + 1)        # length sum
    stats[3 + 2] = 6 * (rank + 1)        # trajectory length terminations
    stats[3 + 5] = 4 * (rank + 1)        # moving obstacle terminations
    metrics = allreduce_episode_stats(stats)
    if rank == 0:
        out.put(metrics)
    dist.destroy_process_group()


def test_episode_stats_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    metrics = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert metrics["episodes"] == 30
    assert abs(metrics["episode_reward_mean"] - 5.0) < 1e-12
    assert abs(metrics["episode_len_mean"] - 20.0) < 1e-12
    assert abs(metrics["termination_reason_trajectory_length_rate"] - 0.6) < 1e-12
    assert abs(metrics["termination_reason_collision_with_moving_obstacle_rate"] - 0.4) < 1e-12
