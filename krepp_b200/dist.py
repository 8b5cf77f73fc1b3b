"""Multi-GPU plumbing for the read-sharded mode (mode A): reads are independent units, the index image is
replicated per GPU, every rank takes one contiguous slice of the read stream and there is NO collective on the data
path.  torch.distributed (NCCL on GPUs, gloo in CPU tests) is only used for the barrier, for max-over-ranks timing and
for gathering per-rank counters."""
from __future__ import annotations

import os


def env_rank_world() -> tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [begin, end) slice of n_items for `rank`; slices of all ranks tile [0, n_items) in order."""
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device=None) -> float:
    """Max of a per-rank scalar (e.g. elapsed seconds); identity when not initialised."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(values, device=None):
    """Element-wise sum of a small list of per-rank counters (e.g. reads, records, algorithmic bytes)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [int(v) for v in values]
    t = torch.tensor([int(v) for v in values], dtype=torch.int64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(x) for x in t.tolist()]


def gather_text(text: str) -> list[str]:
    """Rank-ordered list of every rank's text block (output concatenation in read order); identity when single."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [text]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, text)
    return out
