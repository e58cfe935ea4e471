"""Env-batch sharding across GPUs: envs are fully independent (dmfb.py:127-155 holds no cross-chip
state), so rank r simply owns a contiguous range of global env indices and no collective is needed
on the env path.  The global index (cfg.env_base + local index) keys the on-device RNG, so a sharded
run generates exactly the tasks / draws of the single-GPU run, env for env."""


def shard_range(n_total, rank, world_size):
    """[start, stop) of the envs owned by `rank` when n_total envs are split over world_size ranks."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, rem = divmod(int(n_total), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)
