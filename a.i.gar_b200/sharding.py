"""Env sharding across ranks: envs are independent, so rank r of N owns the contiguous global env
ids [r*E/N, (r+1)*E/N); the Philox key carries the GLOBAL id, which makes every result invariant to N.  The
only collective is the optional all-reduce of episode statistics (src/aigar.py:567-581's reduction)."""


def shard_envs(total_envs, world_size, rank):
    """(first_global_env_id, n_envs) of `rank`; remainders go to the low ranks."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, rem = divmod(int(total_envs), int(world_size))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def allreduce_episode_stats(stats, dist=None):
    """stats: tensor [..., 4] of (sum of masses, max mass, frames, deaths) per agent.  Returns the global
    (mean mass, max mass, frames, deaths); SUM for counts, MAX for the maximum."""
    import torch
    flat = stats.reshape(-1, 4).to(torch.float64)
    sums = torch.stack([flat[:, 0].sum(), flat[:, 2].sum(), flat[:, 3].sum()])
    mx = flat[:, 1].max().reshape(1) if flat.numel() else torch.zeros(1, dtype=torch.float64, device=stats.device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    frames = float(sums[1].item())
    return {"mean_mass": float(sums[0].item()) / max(frames, 1.0), "max_mass": float(mx.item()), "frames": frames,
            "deaths": float(sums[2].item())}
