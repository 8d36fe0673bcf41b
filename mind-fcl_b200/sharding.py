"""Query sharding across ranks (one process per GPU).

Every narrowphase query is independent and reads const geometry
(SURVEY.md 8e), so the batch is split by contiguous query-index range, geometry
is replicated per rank and there is NO collective on the data path.  The only
communication is the optional gather of fixed-size result records to rank 0
(torch.distributed all_gather: NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [begin, end) slice of n queries owned by `rank`; sizes differ by at most 1."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_counts(local_counts, n_total: int, rank: int, world: int, device=None):
    """All ranks contribute their slice of a uint32 per-query result; every rank
    returns the full array (used by tests / small result sets)."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return np.asarray(local_counts)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    max_len = max(e - b for b, e in sizes)
    buf = torch.zeros(max_len, dtype=torch.int64, device=device)
    loc = torch.as_tensor(np.asarray(local_counts).astype(np.int64), device=device)
    buf[: loc.numel()] = loc
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    full = np.empty(n_total, np.uint32)
    for r, (b, e) in enumerate(sizes):
        full[b:e] = out[r][: e - b].cpu().numpy().astype(np.uint32)
    return full
