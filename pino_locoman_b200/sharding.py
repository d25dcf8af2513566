"""Batch sharding across GPUs: instances are independent, so the only exchange is a gather of per-instance results
(SURVEY.md 8e).  One process per GPU; works with any torch.distributed backend (NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_bounds(total, rank, world):
    """Contiguous split of `total` instances: ranks < total % world get one extra instance."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_instance_results(local, total=None, group=None):
    """All-gather a per-instance tensor [B_local, ...] into global instance order [total, ...] on every rank."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    total = local.shape[0] * world if total is None else total
    sizes = [shard_bounds(total, r, world)[1] - shard_bounds(total, r, world)[0] for r in range(world)]
    pad = max(sizes)
    buf = local
    if local.shape[0] < pad:    # all_gather needs equal shapes
        buf = torch.cat([local, local.new_zeros((pad - local.shape[0],) + tuple(local.shape[1:]))], 0)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf.contiguous(), group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], 0)
