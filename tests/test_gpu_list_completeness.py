"""The device's neighbour lists are complete: on MANY different evolved states (every replica of a batch runs on with
its own noise from the bench frame, then from a late production frame of box 11) the forces of the list-based device
path equal the fp64 oracle's, whose lists come from its own, independent cell-list build.  One missed in-cutoff pair
moves a force by many times the 1e-5 tolerance.  (The build walks per-bead spans of a (class, row, x cell) grid with a
half-list clip, clipped row ranges and a bounding-box gate for small classes: this exercises all of them on ~2e6
pairs in cluster, dilute and wall-near geometry.)"""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from enhancershoebox_b200 import datafile, engine as E, protocol as P

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.parametrize("box,snap,tag,first,n_chunks", [(12, "snap_box12_chunk1000.npz", "1000", 1000, 4), (11, "snap_box11.npz", "500", 500, 3)])
def test_lists_complete_on_evolved_states(oracle_lib, pair_tables, oracle_lookup, box, snap, tag, first, n_chunks):
    z = np.load(os.path.join(GOLDEN, snap))
    x, v, t = z["x" + tag].astype(np.float64), z["v" + tag].astype(np.float64), z["t" + tag].astype(np.int32)
    d = datafile.make_shoebox(box, 3, False, seed=7)
    R = 64
    eng = E.ShoeboxBatch(d, n_replicas=R, seeds=np.arange(R) + 31)
    eng.upload_state(x[None], v[None], t[None], replicate=True)
    eng.set_schedule(plans=P.chunk_schedule("genes_stages", n_chunks=first + n_chunks)[first:], observe=True)
    eng.run(0, n_chunks)
    assert np.all(eng.status() == 0)
    xs, vs, ts = eng.state()
    assert not np.array_equal(xs[0], xs[1])                       # the replicas have gone their own ways
    fg, _ = eng.forces("B")
    orc = oracle_lib.Oracle(d, pair_tables, oracle_lookup, seed=1)
    worst = 0.0
    for r in range(R):
        orc.set_state(ts[r], xs[r], vs[r])
        orc.set_pair("B")
        fo, _, st = orc.forces()
        assert st == 0
        worst = max(worst, np.abs(fg[r] - fo).max() / np.abs(fo).max())
    assert worst < TOL, worst
    assert np.all(eng.status() == 0)
    eng.close()
