"""Market sharding across GPUs: markets are independent, so rank r of G owns a
contiguous block of global market ids and the step needs NO collective.  RNG seeds are keyed by the
GLOBAL market id, so results do not depend on G.  Only when one policy batch spans GPUs are the
per-rank observations/rewards all-gathered (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
import numpy as np


def shard_range(num_markets_total, rank, world):
    """Contiguous [lo, hi) block of global market ids owned by `rank` (sizes differ by at most 1)."""
    if not (0 <= rank < world) or num_markets_total < 0:
        raise ValueError("bad rank/world")
    base, rem = divmod(num_markets_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_seeds(base_seed, num_markets_total, rank, world):
    lo, hi = shard_range(num_markets_total, rank, world)
    return (np.arange(lo, hi, dtype=np.uint64) + np.uint64(base_seed))


def shard_slice(global_array, rank, world, axis=0):
    """The slice of a replicated [M_total, ...] action array this rank consumes (no scatter needed)."""
    lo, hi = shard_range(global_array.shape[axis], rank, world)
    idx = [slice(None)] * global_array.ndim
    idx[axis] = slice(lo, hi)
    return global_array[tuple(idx)]


def all_gather_rows(local, num_markets_total, group=None):
    """All-gather per-rank row blocks [M_r, ...] into [M_total, ...] in global market order.
    Uses all_gather_into_tensor when every rank has the same row count, else padded all_gather."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = [shard_range(num_markets_total, r, world) for r in range(world)]
    counts = [hi - lo for lo, hi in sizes]
    if len(set(counts)) == 1:
        out = torch.empty((num_markets_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], 0)
