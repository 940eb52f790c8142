"""Stream sharding across the GPUs of one box.

Streams are independent instances of one program (every piece of state the reference holds in process
globals is per stream here), so the batch shards with NO data-path collective: rank r owns a contiguous
range of streams, its own Executor and its own state.  NCCL is only used for the optional gather of
output PCM onto every rank / rank 0.
"""
from __future__ import annotations


def shard_range(n_streams: int, rank: int, world: int):
    """Contiguous, balanced partition: returns (first_stream, count) of `rank`."""
    base, rem = divmod(n_streams, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def shard_seeds(seeds, rank: int, world: int):
    first, n = shard_range(len(seeds), rank, world)
    return seeds[first:first + n]


def gather_outputs(y_local, n_streams: int, group=None):
    """Optional result gather: every rank receives the full [n_streams, ...] output tensor.
    Works with NCCL (CUDA tensors) and gloo (CPU tensors); shards may differ in size by one stream."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = [shard_range(n_streams, r, world)[1] for r in range(world)]
    pad = max(counts)
    buf = y_local
    if y_local.shape[0] < pad:
        buf = torch.zeros((pad,) + tuple(y_local.shape[1:]), dtype=y_local.dtype, device=y_local.device)
        buf[: y_local.shape[0]] = y_local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf.contiguous(), group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
