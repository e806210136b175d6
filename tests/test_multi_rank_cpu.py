"""N > 1 host logic on CPU (gloo, world_size 2): env sharding by global env id and the one optional collective
(the all-reduce of episode statistics).  The per-rank stepping engine here is the CPU oracle — the CUDA path is
exercised by tests/test_gpu_parity.py::test_shard_invariance on the GPU box."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, total_envs, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import aigar_b200.layout as lay
    from aigar_b200.sharding import shard_envs, allreduce_episode_stats
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, n = shard_envs(total_envs, world, rank)
    cfg = lay.derive_config()
    stats, hashes = [], []
    for i in range(n):
        e = orc.OracleEnv(cfg, seed=21, env_id=first + i)
        e.rollout_random(6, 8, 0)
        b = e.record.players["bot"][0]
        stats.append([b["stat_mass_sum"], b["stat_mass_max"], b["stat_frames"], b["stat_deaths"]])
        hashes.append(int(e.record.header["event_hash"][0]))
    red = allreduce_episode_stats(torch.tensor(stats, dtype=torch.float64), dist)
    gathered = [None] * world
    dist.all_gather_object(gathered, (first, hashes))
    if rank == 0:
        q.put((red, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one_rank():
    import torch.multiprocessing as mp
    import aigar_b200.layout as lay
    from oracle import oracle as orc
    total = 10
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    red, gathered = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process truth
    cfg = lay.derive_config()
    s, mx, fr, hashes = 0.0, 0.0, 0.0, []
    for i in range(total):
        e = orc.OracleEnv(cfg, seed=21, env_id=i)
        e.rollout_random(6, 8, 0)
        b = e.record.players["bot"][0]
        s, mx, fr = s + float(b["stat_mass_sum"]), max(mx, float(b["stat_mass_max"])), fr + float(b["stat_frames"])
        hashes.append(int(e.record.header["event_hash"][0]))
    assert red["frames"] == fr and red["max_mass"] == mx and red["mean_mass"] == pytest.approx(s / fr, rel=1e-12)
    merged = [h for _, hs in sorted(gathered) for h in hs]
    assert merged == hashes  # results are invariant to the shard count: the Philox key carries the global env id
