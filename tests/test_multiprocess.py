"""N>1 path on CPU: two gloo ranks each step their shard of the replicas (host-emulated kernels);
the gathered result must equal a single-process run over all replicas (sharding invariance)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pednstream_b200.parallel import gather_replica_values, max_over_ranks, shard_replicas

TOTAL, STEPS = 5, 25


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_env(replicas, base):
    import ctypes
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tests", "emu"))
    import build_emu
    from pednstream_b200 import _native
    from pednstream_b200.rl import BatchedPedNetEnv
    lib = _native._declare(ctypes.CDLL(build_emu.build()))
    env = BatchedPedNetEnv("nine_intersections", replicas=replicas, obs_mode="option3", seed=77,
                           replica_base=base, _lib=lib, _emulation=True)
    rs = np.random.RandomState(3)
    acts = rs.uniform(0, 4, size=(STEPS, TOTAL, env.n_act)).astype(np.float32)
    for k in range(STEPS):
        env.step(torch.from_numpy(acts[k, base:base + replicas].copy()))
    return env.cumulative_reward.clone(), env.obs.clone()


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    count, base = shard_replicas(TOTAL, world, rank)
    rew, obs = _run_env(count, base)
    all_rew = gather_replica_values(rew, TOTAL)
    all_obs = gather_replica_values(obs, TOTAL)
    slowest = max_over_ranks([float(rank + 1), 2.0])
    if rank == 0:
        torch.save({"rew": all_rew, "obs": all_obs, "slowest": slowest}, os.path.join(out_dir, "gathered.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_replicas_covers_range():
    for total, world in ((8192, 8), (5, 2), (3, 4), (7, 7)):
        spans = [shard_replicas(total, world, r) for r in range(world)]
        assert sum(c for c, _ in spans) == total
        nxt = 0
        for c, first in spans:
            assert first == nxt
            nxt += c


def test_two_ranks_equal_single_process(tmp_path, emu_lib):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = torch.load(os.path.join(tmp_path, "gathered.pt"))
    rew, obs = _run_env(TOTAL, 0)
    assert torch.equal(got["rew"], rew)
    assert torch.equal(got["obs"], obs)
    assert got["slowest"] == [2.0, 2.0]
