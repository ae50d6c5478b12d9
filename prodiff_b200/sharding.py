"""Batch sharding of the sampler across the GPUs of one box.

Utterances are independent (no cross-row op anywhere between `prodiff.py:136` and `wavenet.py:123`), so the batch is
split into contiguous row blocks, one per rank, weights are replicated, and the hot path has NO collective.  The only
communication is an optional all-gather of the finished mels, outside the sampling loop.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_rows: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced [start, end) row ranges; the first ``n_rows % world_size`` ranks get one extra row."""
    if n_rows < 0 or world_size < 1:
        raise ValueError("n_rows >= 0 and world_size >= 1 required")
    base, extra = divmod(n_rows, world_size)
    out, s = [], 0
    for r in range(world_size):
        e = s + base + (1 if r < extra else 0)
        out.append((s, e))
        s = e
    return out


def sample_sharded(sample_fn: Callable[[torch.Tensor], torch.Tensor], cond: torch.Tensor, gather: bool = False,
                   group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Run ``sample_fn`` (cond_shard [b,T,H] -> mel [b,T,M]) on this rank's rows of ``cond`` [B,T,H].

    Every rank passes the same full ``cond`` (or at least its own rows filled in).  Returns this rank's mel shard, or
    with ``gather=True`` the full [B,T,M] on every rank (one all-gather of padded shards, after sampling).
    """
    if not (dist.is_available() and dist.is_initialized()):
        return sample_fn(cond)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bounds = shard_bounds(cond.shape[0], world)
    s, e = bounds[rank]
    mel = sample_fn(cond[s:e]) if e > s else cond.new_zeros((0, cond.shape[1], 0))
    if not gather:
        return mel
    # shards may differ by one row: pad to the largest, gather, trim
    rows = max(b - a for a, b in bounds)
    M = torch.tensor([mel.shape[2] if e > s else 0], device=cond.device)
    dist.all_reduce(M, op=dist.ReduceOp.MAX, group=group)
    M = int(M.item())
    padded = cond.new_zeros((rows, cond.shape[1], M))
    if e > s:
        padded[: e - s] = mel
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded.contiguous(), group=group)
    return torch.cat([o[: b - a] for o, (a, b) in zip(out, bounds)], dim=0)
