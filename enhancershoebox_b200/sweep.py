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
    if device is None and "nccl" in str(dist.get_backend(group)):
        device = torch.device("cuda", torch.cuda.current_device())      # NCCL moves device tensors only
    if device is not None:
        t = t.to(device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    return np.concatenate([p.cpu().numpy()[:c] for p, c in zip(parts, counts)], axis=0)


def sweep_groups(promoters, activations, thresholds, repeats: int, group: int = 5):
    """(promoter, threshold, activation, first_repeat) of every `group`-run block of the sweep, in the aggregator's
    loop order (VisitorGeneMaps/dist_genestages_grouped.py:45-54, 70-106): one output row each."""
    out = []
    for p in promoters:
        for a in activations:
            for t in thresholds:
                for g in range(repeats // group):
                    out.append((p, t, a, 1 + g * group))
    return out


def run_sweep(box: int, promoters=(1, 2, 3), activations=(30,), thresholds=(10, 20, 30, 40, 50, 60, 70, 80, 90, 100),
              repeats: int = 5, n_chunks: int | None = None, group: int = 5, batch_groups: int = 118, device: int = 0,
              seed_base: int = 0, retries: int = 3, progress=None):
    """The GENE_STAGES sweep (VisitorGeneMaps/run_singleCond_parallel.sh + dist_genestages_grouped.py) without the
    filesystem in between: the 5-run groups are dealt to the ranks (one process per GPU, no collective on the step
    path), every rank runs its groups as batches of replicas, reduces each group's geneTrack records to the summary
    row ON the device (the frames never leave HBM) and the rows are all-gathered at the end.

    Returns (groups, rows): groups = sweep_groups(...) and rows [len(groups), 4] = S5PInt, S2PInt, Contact,
    DistActivation on every rank.  A replica that dies (status != 0; LAMMPS would have aborted that run) is redone
    with a fresh seed, as the reference's launcher does (run_parallel_shoebox.sh:1-14)."""
    import ctypes as C

    import torch.distributed as dist

    from . import datafile, engine as E, protocol as P
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    groups = sweep_groups(promoters, activations, thresholds, repeats, group)
    lo, hi = shard_range(len(groups), rank, world)
    mine = list(range(lo, hi))
    rows = np.full((hi - lo, 4), np.nan)
    total = P.N_RUNS if n_chunks is None else n_chunks
    lib = E.load_library()
    attempt = {g: 0 for g in mine}
    for p in promoters:
        pending = [g for g in mine if groups[g][0] == p]
        while pending:
            now, pending = pending[:batch_groups], pending[batch_groups:]
            runs = [(g, k) for g in now for k in range(group)]
            seeds = np.array([seed_base + 1_000_003 * attempt[g] + (g * group + k) + 1 for g, k in runs], dtype=np.uint64)
            datas = [datafile.make_shoebox(box, p, False, seed=int(s)) for s in seeds]
            eng = E.ShoeboxBatch(datas, seeds=seeds, thresholds=[groups[g][1] for g, _ in runs],
                                 activations=[groups[g][2] for g, _ in runs], device=device)
            eng.set_schedule(total, observe=True)
            eng.run()
            ok = (eng.status() == 0).reshape(len(now), group).all(axis=1)
            out = np.empty((len(now), 4))
            rc = lib.esb_aggregate_grouped(device, C.c_void_p(eng.device_ptr(2)), 1, len(runs), total, group, 250.0, P.T_INDUCTION_ON,
                                           out.ctypes.data_as(C.POINTER(C.c_double)), None)
            eng.close()
            if rc != 0:
                raise E.EsbError(f"esb_aggregate_grouped failed with {rc}")
            for q, g in enumerate(now):
                if ok[q]:
                    rows[g - lo] = out[q]
                elif attempt[g] < retries:
                    attempt[g] += 1
                    pending.append(g)
            if progress:
                progress(rank, p, len(now), len(pending))
    gather_dev = None
    if dist.is_available() and dist.is_initialized() and "nccl" in str(dist.get_backend()):
        import torch
        gather_dev = torch.device("cuda", device)
    return groups, gather_rows(rows, len(groups), device=gather_dev)


def summary_csv(groups, rows) -> str:
    """summary_contact_grouped_*.txt as `df_summary.to_csv` writes it (VisitorGeneMaps/dist_genestages_grouped.py:101-111)."""
    import pandas as pd
    cols = ['Promoter', 'Threshold', 'Activation', 'S5PInt', 'S2PInt', 'Contact', 'DistActivation']
    df = pd.DataFrame(columns=cols)
    parts = [pd.DataFrame(data=np.array([[p, t, a, r[0], r[1], r[2], r[3]]], dtype=object), columns=cols) for (p, t, a, _), r in zip(groups, rows)]
    if parts:
        df = pd.concat([df] + parts, ignore_index=True)
    return df.to_csv()


def summary_cpp_csv(groups, rows) -> str:
    """summary_contact_grouped_*_cpp.txt as VisitorGeneMaps/dist_genestages_grouped_cpp writes it (binary only; golden
    VisitorGeneMaps/summary_contact_grouped_Thresholds10-100_cpp.txt): no index column, `ostream << double` with the
    default precision (6 significant digits = printf %g), and 0 where the Python aggregator leaves the mean of an empty
    list (DistActivation of a group without activation) as NaN."""
    lines = ["Promoter,Threshold,Activation,S5PInt,S2PInt,Contact,DistActivation"]
    for (p, t, a, _), r in zip(groups, rows):
        vals = ["%g" % (0.0 if np.isnan(v) else v) for v in r]
        lines.append(",".join([str(int(p)), str(int(t)), str(int(a))] + vals))
    return "\n".join(lines) + "\n"
