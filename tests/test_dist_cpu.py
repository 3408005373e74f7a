"""CPU, world_size 2, gloo: the N>1 host logic -- contiguous env sharding and the all-gather of the per-rank
episode-statistics vector (the framework's only collective; NCCL over NVLink in bench.py, gloo here)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gym_collision_avoidance_b200 import abi
from gym_collision_avoidance_b200.envs.vec_env import gather_stats, shard_envs, stats_vector


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, n = shard_envs(1001, rank, world)
    stats = {name: float((rank + 1) * (i + 1)) for i, name in enumerate(abi.STAT_NAMES)}
    stats["episodes"] = float(n)
    tot, per_rank = gather_stats(stats)
    out[rank] = (first, n, tot, per_rank.tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    (f0, n0, tot0, pr0), (f1, n1, tot1, pr1) = out[0], out[1]
    assert (f0, n0, f1, n1) == (0, 501, 501, 500)
    assert tot0 == tot1 and pr0 == pr1
    assert tot0["episodes"] == 1001.0
    for i, name in enumerate(abi.STAT_NAMES[1:], start=1):
        assert tot0[name] == 3.0 * (i + 1)
    assert np.asarray(pr0).shape == (2, len(abi.STAT_NAMES))


def test_shard_envs_partitions_exactly():
    for total in (1, 7, 8192, 16385):
        for world in (1, 2, 4, 8):
            blocks = [shard_envs(total, r, world) for r in range(world)]
            assert sum(n for _, n in blocks) == total
            pos = 0
            for first, n in blocks:
                assert first == pos
                pos += n
    v = stats_vector({n: i for i, n in enumerate(abi.STAT_NAMES)})
    assert v.tolist() == list(map(float, range(len(abi.STAT_NAMES))))
