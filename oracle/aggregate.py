"""numpy restatement of the reference aggregation (TEST INFRASTRUCTURE, see esb_oracle.c).

Follows ``VisitorGeneMaps/dist_genestages_grouped.py:70-106`` (groups of 5 runs, in the order the
runs are handed in) and ``VisitorGeneMaps/dist_genestages_all.py:66-85`` (one row per run).
Pinned against the reference script itself: tests/golden/aggregate_golden.npz (made by
tests/golden/make_aggregate_golden.py, which runs the unmodified script in this container)."""
from __future__ import annotations

import warnings

import numpy as np


def _mean(values):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return float(np.mean(values))


def grouped(frames: np.ndarray, group: int = 5, contact_dist: float = 250, t_skip: int = 150) -> np.ndarray:
    """frames [n_runs, n_frames, 9] -> rows [n_runs // group, 4] = S5PInt, S2PInt, Contact, DistActivation."""
    rows = []
    s5p_group, active_group, contact_group, dist_group = [], [], [], []
    for i_r, data in enumerate(frames, start=1):
        d_rp = [float(x[4]) for x in data]
        gene_state = [float(x[8]) for x in data]
        s5p_promoter = [float(x[6]) for x in data]
        I = [x for x in range(1, len(gene_state)) if int(gene_state[x]) == 2 and int(gene_state[x - 1]) != 2]
        J = [x for x in range(1, len(gene_state)) if int(gene_state[x]) == 2]
        s5p_group.extend(s5p_promoter)
        active_group.append(100 * len(J) / (len(gene_state) - t_skip))
        contact_group.append(100 * sum(np.array(d_rp) < contact_dist) / len(d_rp))
        dist_group.extend([d_rp[x] for x in I])
        if i_r % group == 0:
            rows.append([_mean(s5p_group), _mean(active_group), _mean(contact_group), _mean(dist_group)])
            s5p_group, active_group, contact_group, dist_group = [], [], [], []
    return np.array(rows).reshape(-1, 4)


def per_run_percentile(frames: np.ndarray, q: float = 5, contact_dist: float = 250, t_skip: int = 150) -> np.ndarray:
    """frames [n_runs, n_frames, 9] -> rows [n_runs, 6] = mean s5p, S2PInt, Contact, DistActivation,
    q-percentile of d_rg[t_skip:], of d_rp[t_skip:]
    (``VisitorGeneMaps/dist_genestages_grouped_5percentile_10xaveraging.py:72-89``)."""
    rows = []
    for data in frames:
        d_rp = [float(x[4]) for x in data]
        d_rg = [float(x[5]) for x in data]
        gene_state = [float(x[8]) for x in data]
        s5p_promoter = [float(x[6]) for x in data]
        I = [x for x in range(1, len(gene_state)) if int(gene_state[x]) == 2 and int(gene_state[x - 1]) != 2]
        J = [x for x in range(1, len(gene_state)) if int(gene_state[x]) == 2]
        rows.append([float(np.mean(s5p_promoter)), 100 * len(J) / (len(gene_state) - t_skip),
                     100 * sum(np.array(d_rp) < contact_dist) / len(d_rp),
                     float(np.mean([d_rp[x] for x in I])) if len(I) > 0 else np.nan,
                     float(np.percentile(d_rg[t_skip:], q)), float(np.percentile(d_rp[t_skip:], q))])
    return np.array(rows).reshape(-1, 6)


def percentile_10x(frames: np.ndarray, group: int = 10, **kw):
    """(labels, rows [ceil(n_runs / group), 6]): np.nanmean over consecutive groups of `group` runs and over the
    remainder (``…_5percentile_10xaveraging.py:94-141``)."""
    per_run = per_run_percentile(frames, **kw)
    labels, rows = [], []
    for lo in range(0, len(per_run), group):
        hi = min(lo + group, len(per_run))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            rows.append([float(np.nanmean(list(per_run[lo:hi, c]))) for c in range(6)])
        labels.append(f"{lo + 1}-{hi}")
    return labels, np.array(rows).reshape(-1, 6)
