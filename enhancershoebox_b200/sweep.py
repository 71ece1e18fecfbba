"""Replica / condition sharding across the GPUs of one box and the end-of-run gather.

The reference's only parallelism is independent OS processes, one per repeat, throttled by a FIFO
semaphore (``run_parallel_shoebox.sh:17-41,58-76``; ``VisitorGeneMaps/run_singleCond_parallel.sh``);
results meet on the filesystem.  Here replicas are dealt statically to ranks (one process per GPU),
nothing crosses ranks on the step path, and at the end the small per-replica tables are gathered
with one `all_gather` (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of replica indices [lo, hi) owned by `rank` (blocks differ by at most 1)."""
    if not (0 <= rank < world):
        raise ValueError("rank outside world")
    base, extra = divmod(n_total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_counts(n_total: int, world: int) -> list[int]:
    return [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]


def sweep_conditions(promoters=(1, 2, 3), activations=(1, 5, 10, 15, 20, 25, 30, 50, 75, 100),
                     thresholds=(10, 20, 30, 40, 50, 60, 70, 75, 80, 90, 100), repeats=100):
    """The (promoter, threshold, activation, repeat) grid of the sweep launcher
    (VisitorGeneMaps/run_singleCond_parallel.sh:40-61) in the aggregator's loop order
    (VisitorGeneMaps/dist_genestages_grouped.py:45-54)."""
    out = []
    for p in promoters:
        for a in activations:
            for t in thresholds:
                for r in range(1, repeats + 1):
                    out.append((p, t, a, r))
    return out


def gather_rows(local: np.ndarray, n_total: int, group=None, device=None) -> np.ndarray:
    """All-gather per-replica rows [n_local, ...] into [n_total, ...] on every rank.

    Uses torch.distributed when it is initialised (any backend), otherwise returns `local`.
    Ranks may own different counts (shard_range); rows are padded to the largest shard for the
    collective and trimmed afterwards.
    """
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        assert local.shape[0] == n_total
        return local
    world = dist.get_world_size(group)
    counts = shard_counts(n_total, world)
    rank = dist.get_rank(group)
    assert local.shape[0] == counts[rank], (local.shape, counts, rank)
    width = max(counts)
    t = torch.zeros((width,) + tuple(local.shape[1:]), dtype=torch.float64)
    t[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64))
    if device is not None:
        t = t.to(device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    return np.concatenate([p.cpu().numpy()[:c] for p, c in zip(parts, counts)], axis=0)
