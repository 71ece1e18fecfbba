"""Oracle parity on EVERY compiled bead capacity the bench and BASELINE.json's configs use: box 9 (1050 beads ->
kernels<1056>), box 10 (1172 -> <1184>), box 12 (1400 -> <1408>, the benchmarked instantiation, incl. the bench's own
chunk-1000 frame); box 11 (<1312>) is covered by test_gpu_forces.py / test_gpu_trajectory.py, the generic and the
actin (<1600>) instantiations by test_gpu_forces.py::test_generic_size_path and test_gpu_actin.py.
Snapshots: tests/golden/snap_box{9,10,12}.npz (oracle runs from make_shoebox(box, 3, seed=7),
tests/golden/make_oracle_snapshots.py).  Tolerances as in test_gpu_forces.py / test_gpu_trajectory.py."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from enhancershoebox_b200 import datafile, engine as E, protocol as P

pytestmark = pytest.mark.gpu
TOL = 1e-5
N_PAD = {9: 1056, 10: 1184, 12: 1408}


def _case(box, tag):
    z = np.load(os.path.join(GOLDEN, "snap_box12_chunk1000.npz" if tag == "1000" else f"snap_box{box}.npz"))
    return z["x" + tag].astype(np.float64), z["v" + tag].astype(np.float64), z["t" + tag].astype(np.int32)


@pytest.fixture(scope="module", params=[9, 10, 12])
def system(request, oracle_lib, pair_tables, oracle_lookup):
    box = request.param
    d = datafile.make_shoebox(box, 3, False, seed=7)
    eng = E.ShoeboxBatch(d, n_replicas=3, seeds=[4242, 4243, 4244])
    assert eng.n_pad == N_PAD[box]
    orc = oracle_lib.Oracle(d, pair_tables, oracle_lookup, seed=4242)
    yield box, d, eng, orc
    eng.close()


@pytest.mark.parametrize("tag,pair", [("9", "soft"), ("14", "A"), ("500", "B"), ("1000", "B")])
@pytest.mark.parametrize("use_arrays", [False, True])
def test_forces_and_energies(system, oracle_lib, tag, pair, use_arrays):
    box, d, eng, orc = system
    if tag == "1000" and box != 12:
        pytest.skip("the bench frame is a box-12 state")
    x, v, t = _case(box, tag)
    orc.set_state(t, x, v)
    orc.set_pair(pair)
    if pair == "soft":
        orc.L.esbo_pair_soft(orc.h, 54.0, oracle_lib._dp(orc.soft_cut))
    fo, eo, st = orc.forces()
    assert st == 0
    eng.upload_state(np.stack([x] * 3), np.stack([v] * 3), np.stack([t] * 3))
    fg, eg = eng.forces(pair, soft_a=54.0, use_arrays=use_arrays)
    assert np.abs(fg[0] - fo).max() / np.abs(fo).max() < TOL
    assert np.array_equal(fg[0], fg[1]) and np.array_equal(fg[0], fg[2])
    assert np.allclose(eg[0][:3], eo[:3], rtol=2e-6, atol=1e-3)
    fo2, eo2, _ = orc.forces(with_post=True, with_noise=False)
    orc.set_seed(4242, 5)
    fo3, _, _ = orc.forces(with_post=True, with_noise=True)
    fg3, _ = eng.forces(pair, soft_a=54.0, with_post=True, eval_index=5, use_arrays=use_arrays)
    assert np.abs(fg3[0] - fo3).max() / np.abs(fo3).max() < TOL
    assert np.all(eng.status() == 0)


@pytest.mark.parametrize("tag", ["500", "1000"])
def test_short_trajectory_matches_oracle(system, oracle_lib, pair_tables, oracle_lookup, tag):
    """25 zero-noise steps (drag only) of the step kernel against the fp64 oracle; same stated tolerance as
    test_gpu_trajectory.py: median |dx| < 1e-5, 99 % of the beads within 1e-4 (bead diameter = 1)."""
    box, d, _, _ = system
    if tag == "1000" and box != 12:
        pytest.skip("the bench frame is a box-12 state")
    x, v, t = _case(box, tag)
    first = int(tag)
    plans = [P.ChunkPlan(p.index, p.pair, p.ramp, p.adapt_soft, p.adapt_bond1, p.adapt_bond2, 25)
             for p in P.chunk_schedule("genes_stages", n_chunks=first + 1)[first:]]
    results = []
    for R in (2, 300):                                   # 2 replicas: cluster variant; 300: the full-batch kernel the bench runs
        eng = E.ShoeboxBatch(d, n_replicas=R, seeds=[99] * R, langevin_mode=2)
        eng.upload_state(x[None], v[None], t[None], replicate=True)
        eng.set_schedule(plans=plans, observe=False)
        eng.run(0, 1)
        xg, vg, _ = eng.state()
        assert np.all(eng.status() == 0)
        results.append((xg[0], vg[0]))
        assert np.array_equal(xg[0], xg[-1])
        eng.close()
    assert np.array_equal(results[0][0], results[1][0]) and np.array_equal(results[0][1], results[1][1])   # every kernel variant: same bits
    orc = oracle_lib.Oracle(d, pair_tables, oracle_lookup, seed=99)
    orc.set_langevin(2)
    orc.set_state(t, x, v)
    orc.apply_plan(plans[0])
    assert orc.run(25) == 0
    xo, vo, _, _ = orc.state()
    dx = np.abs(results[1][0] - xo).max(axis=1)
    assert np.median(dx) < 1e-5 and np.mean(dx < 1e-4) > 0.99 and dx.max() < 0.05, (np.median(dx), np.mean(dx < 1e-4), dx.max())
    assert np.median(np.abs(results[1][1] - vo).max(axis=1)) < 1e-4
