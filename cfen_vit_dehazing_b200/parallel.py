"""Image sharding of a batch across the GPUs of one box.

After ActNorm is initialised every op of the generator is per image, so the batch partitions
contiguously, weights are replicated and there is NO collective on the data path (the reference's own
multi-GPU mechanism, ``nn.DataParallel`` at v3.py:81, scatters/gathers the same way).  This module is
pure arithmetic plus an optional output gather, so it can be tested with the gloo backend on CPU.
"""
import torch

__all__ = ["shard_range", "shard_sizes", "shard_batch", "gather_outputs"]


def shard_sizes(batch, world_size):
    """Images per rank: contiguous, sizes differ by at most one, earlier ranks take the remainder
    (``torch.chunk``-like for divisible batches, but never yields fewer than ``world_size`` shards)."""
    if batch < 0 or world_size < 1:
        raise ValueError("batch must be >= 0 and world_size >= 1")
    base, rem = divmod(batch, world_size)
    return [base + (1 if r < rem else 0) for r in range(world_size)]


def shard_range(batch, world_size, rank):
    """[start, stop) of the images rank ``rank`` processes."""
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    sizes = shard_sizes(batch, world_size)
    start = sum(sizes[:rank])
    return start, start + sizes[rank]


def shard_batch(x, world_size, rank):
    """This rank's images of a [B, ...] tensor (a view, no copy)."""
    a, b = shard_range(x.shape[0], world_size, rank)
    return x[a:b]


def gather_outputs(outs, batch, group=None):
    """Optional: assemble the full-batch [xr, xs, xd] on every rank from per-rank shards (outside any timed
    region; the forward path itself never communicates).  Ragged shards are padded to the largest one."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = shard_sizes(batch, world)
    big = max(sizes)
    full = []
    for o in outs:
        pad = o.new_zeros((big,) + tuple(o.shape[1:]))
        pad[:o.shape[0]] = o
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        full.append(torch.cat([p[:n] for p, n in zip(parts, sizes)], 0))
    return full
