"""Clip / video sharding across ranks and the one collective of the path.

Clips are independent in eval, so N GPUs = N processes that each run the whole model on a slice of the
batch; the only exchange is an all-gather of per-clip scores, mirroring `dist.all_gather(preds)` in the reference
(trainer_ddp.py:259-267).  Unlike the reference -- which silently assumes equal-length shards -- the last shard is
padded and the padding is trimmed after the gather."""
import torch


def shard_bounds(n_items, world, rank):
    """Contiguous shard [lo, hi) of rank; shards differ by at most one item, earlier ranks get the larger ones."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def padded_shard_size(n_items, world):
    return (n_items + world - 1) // world


def all_gather_scores(local_scores, n_items, group=None):
    """local_scores: 1-D tensor with this rank's shard (shard_bounds order).  Returns the [n_items] tensor in global
    order on every rank.  Works with NCCL (CUDA tensors) and gloo (CPU tensors); single process = identity."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        if local_scores.numel() != n_items:
            raise ValueError("single-process gather needs the full score vector")
        return local_scores
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(n_items, world, rank)
    if local_scores.numel() != hi - lo:
        raise ValueError(f"rank {rank} holds {local_scores.numel()} scores, its shard has {hi - lo}")
    pad = padded_shard_size(n_items, world)
    send = torch.zeros(pad, dtype=local_scores.dtype, device=local_scores.device)
    send[:hi - lo] = local_scores
    recv = torch.empty(world * pad, dtype=local_scores.dtype, device=local_scores.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    parts = []
    for r in range(world):
        rlo, rhi = shard_bounds(n_items, world, r)
        parts.append(recv[r * pad:r * pad + (rhi - rlo)])
    return torch.cat(parts)
