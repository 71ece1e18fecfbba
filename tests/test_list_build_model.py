"""Geometry of the device's list build (numpy model in oracle/list_build_model.py, float32 like the kernel) against a
brute-force pair search: every pair within cut + skin is listed, exactly once, nothing beyond it -- on random systems and on
the configurations the GPU tests cannot enumerate: beads exactly on row and cell boundaries, on the walls, classes of
1 / 32 / 33 beads (the small-class gate switches at 32), empty classes, the bench frame itself."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import list_build_model as M

CLASS_OF = np.array([0, 0, 0, 1, 2, 3, 5, 5, 4, 0, 0, 0])      # k_class_of, esb_api.cu


def _cut_matrix(rng, nt=12):
    """Every pair of real types interacts (LAMMPS insists on all pair coefficients: run_single_shoebox.py:800-874 sets them all)."""
    c = rng.choice([0.33, 0.71, 1.09, 1.42, 1.77, 2.73], size=(nt, nt))
    c = np.maximum(c, c.T)
    c[0] = c[:, 0] = 0.0
    return c


def _check(x, types, cut, lo, hi, **grid_kw):
    n_pad = (len(x) + 31) & ~31
    g = M.Grid(np.asarray(lo, float), np.asarray(hi, float), max(n_pad, 64), **grid_kw)
    pairs, tests = M.build_half_list(x, types, CLASS_OF, cut, g)
    sure, maybe = M.brute_force(x, types, cut)
    got = set(pairs)
    assert len(pairs) == len(got), "a pair is listed twice"
    assert sure <= got, sorted(sure - got)[:5]
    assert got <= (sure | maybe), sorted(got - sure - maybe)[:5]
    return len(got), tests


@pytest.mark.parametrize("seed", range(6))
def test_random_systems(seed):
    rng = np.random.default_rng(seed)
    lo, hi = np.array([-12.0, -7.0, -4.5]), np.array([12.0, 7.0, 4.5])
    n = 500
    x = rng.uniform(lo, hi, size=(n, 3))
    x[:150] = rng.normal([3.0, 1.0, -0.5], 1.2, size=(150, 3)).clip(lo + 1e-3, hi - 1e-3)      # a dense cluster
    types = rng.choice([1, 4, 5, 6, 8, 9], size=n, p=[0.25, 0.25, 0.15, 0.01, 0.3, 0.04])
    types[:150] = 8
    npairs, tests = _check(x, types, _cut_matrix(rng), lo, hi)
    assert npairs > 500 and tests < 20 * npairs


def test_beads_on_boundaries_and_walls():
    """Coordinates that are exact multiples of the row edge and of the x-cell edge, and beads on the box faces."""
    rng = np.random.default_rng(11)
    lo, hi = np.array([-12.0, -7.0, -4.2]), np.array([12.0, 7.0, 4.2])
    g = M.Grid(lo, hi, 512)
    ys = lo[1] + g.edge * np.arange(g.nry + 1)
    zs = lo[2] + g.edge * np.arange(g.nrz + 1)
    xs = lo[0] + (hi[0] - lo[0]) / g.ncx * np.arange(g.ncx + 1)
    n = 400
    x = np.stack([rng.choice(xs, n), rng.choice(ys, n), rng.choice(zs, n)], axis=1)
    x += rng.choice([0.0, 0.0, 1e-6, -1e-6, 0.3, -0.3], size=(n, 3))
    x = np.clip(x, lo, hi)
    x[:20, 2] = hi[2]; x[20:40, 1] = lo[1]; x[40:60, 0] = hi[0]
    types = rng.choice([1, 4, 5, 8], size=n)
    _check(x, types, _cut_matrix(rng), lo, hi)


@pytest.mark.parametrize("small", [1, 32, 33])
def test_small_and_empty_classes(small):
    """The gene-body class (types 6, 7) with 1, 32 and 33 beads: one gated span up to 32, row spans beyond; actin absent."""
    rng = np.random.default_rng(small)
    lo, hi = np.array([-9.0, -7.0, -4.5]), np.array([9.0, 7.0, 4.5])
    n = 300
    x = rng.uniform(lo, hi, size=(n, 3))
    types = rng.choice([1, 4, 5, 8], size=n)
    types[:small] = 6
    x[:small] = rng.normal([0.0, 0.0, 0.0], 0.8, size=(small, 3))
    cut = _cut_matrix(rng)
    cut[6, :] = cut[:, 6] = 2.73                                   # long cutoffs towards the small class
    _check(x, types, cut, lo, hi)


def test_bench_frame():
    """The bench's own chunk-1000 frame of box 12 with the genes/stages cutoffs of phase B (effective cutoffs as the kernel
    trims them are shorter; the untrimmed ones are the harder case)."""
    from enhancershoebox_b200 import protocol as P
    z = np.load(os.path.join(GOLDEN, "snap_box12_chunk1000.npz"))
    x, t = z["x1000"].astype(np.float32), z["t1000"].astype(np.int64)
    pt = P.pair_tables("genes_stages")
    m = pt.maps["B"]
    cut = np.zeros((12, 12))
    for a in range(12):
        for b in range(12):
            if m[a, b] >= 0:
                cut[a, b] = pt.specs[m[a, b]][1]
    sel = np.arange(0, len(x), 2)                                  # every other bead: the pure-Python model is slow
    npairs, tests = _check(x[sel], t[sel], cut, [-12.0, -7.0, -4.5], [12.0, 7.0, 4.5])
    assert npairs > 3000
