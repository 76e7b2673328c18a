"""Multi-GPU plumbing: how a batch is split over ranks and how the GP scalars come back.

The path shards trivially (every matrix / GP evaluation is independent): contiguous
ranges of the batch, one process per GPU, no collective on the data path.  The only exchange is the
final gather of one scalar per evaluation (or one checksum per rank), done with torch.distributed
(NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations


def shard_bounds(batch: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous shard [lo, hi) of `batch` units for `rank`: ceil(batch / world) per rank."""
    if world < 1 or not (0 <= rank < world) or batch < 0:
        raise ValueError((batch, world, rank))
    per = -(-batch // world)
    lo = min(batch, rank * per)
    return lo, min(batch, lo + per)


def gather_shards(local, batch: int, group=None):
    """all_gather of per-rank result slices (1-D tensors, ragged last shard allowed) -> full tensor on
    every rank, in batch order.  Works on CPU tensors with gloo and CUDA tensors with NCCL."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    per = -(-batch // world)
    padded = torch.zeros(per, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    return torch.cat(out)[:batch]
