"""Multi-GPU: the sampling batch shards across ranks; nothing is exchanged during the loop.

Every op of the denoiser reduces only within one sample, so rank r runs the
identical loop on rows [lo, hi) of the global batch with replicated weights, and ONE all-gather of the
fp32 results reassembles x_0 (it replaces the pickle all_gather of mogen/apis/test.py:131-163).  No
collective is invented on the data path.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous, balanced slice of n rows for `rank` (first n % world ranks get one extra row)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_rows(local, n_total, group=None):
    """All-gather row shards (dim 0) produced with `shard_range` into the full [n_total, ...] tensor on
    every rank.  One collective: all_gather_into_tensor when the shards are equal, otherwise the padded
    form (max shard) followed by a local un-pad."""
    if not (dist.is_available() and dist.is_initialized()):
        assert local.shape[0] == n_total
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    counts = [hi - lo for lo, hi in sizes]
    mx = max(counts)
    local = local.contiguous()
    if min(counts) == mx:
        out = local.new_empty((n_total,) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    padded = local.new_zeros((mx,) + tuple(local.shape[1:]))
    padded[: local.shape[0]] = local
    buf = local.new_empty((world * mx,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(buf, padded, group=group)
    return torch.cat([buf[r * mx: r * mx + counts[r]] for r in range(world)], 0)


def sample_sharded(engine, tables, x_T_local, n_total, group=None):
    """Run the sampler on this rank's rows and all-gather x_0 (device tensors)."""
    x0_local = engine.sample(tables, x_T_local)
    return gather_rows(x0_local, n_total, group)
