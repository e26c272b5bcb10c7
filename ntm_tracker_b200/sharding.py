"""Multi-GPU sharding of the sequence batch.

Sequences are independent (they share only the read-only weights and the tiled
initial state, ntm_cell.py:296,301,306), so the path shards by contiguous ranges
of the batch with NO data-path collective: one process per GPU, weights
replicated, each rank runs the same persistent kernel on its own sequences.  The
only cross-rank operations are harness-level: a barrier and a max-over-ranks of
the device time.  (The reference has no multi-device code at all.)
"""
import torch
import torch.distributed as dist


def shard_range(batch, world_size, rank):
    """Contiguous [lo, hi) of a batch of `batch` sequences owned by `rank`
    (sizes differ by at most one; every sequence is owned exactly once)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    return rank * batch // world_size, (rank + 1) * batch // world_size


def max_over_ranks(value, device=None):
    """Max of a host scalar over all ranks (timing reduction); identity when not distributed."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_batch(local, batch, device=None):
    """All-gather per-rank results [B_local, ...] back into batch order [B, ...]
    (used by tests / callers that want the full result on every rank; the
    hot path itself never needs it)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(batch, world, r)[1] - shard_range(batch, world, r)[0] for r in range(world)]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return torch.cat([o[:n] for o, n in zip(outs, sizes)], 0)
