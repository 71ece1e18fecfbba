"""Level 1 of the north-star checks: single-evaluation forces and energies of the CUDA path against
the fp64 oracle on identical (fp32-representable) coordinates.  Tolerance: 1e-5 of the largest force
component (BASELINE.json north_star: "within 1e-5 relative in fp32"); measured ~3e-7."""
import numpy as np
import pytest

from enhancershoebox_b200 import engine as E

pytestmark = pytest.mark.gpu
TOL = 1e-5
PHASES = (("9", "soft"), ("14", "A"), ("20", "B"), ("500", "B"), ("1999", "B"))


@pytest.fixture(scope="module")
def eng(box11):
    e = E.ShoeboxBatch(box11, n_replicas=3, seeds=[4242, 4243, 4244])
    yield e
    e.close()


@pytest.fixture(scope="module")
def orc(oracle_lib, box11, pair_tables, oracle_lookup):
    return oracle_lib.Oracle(box11, pair_tables, oracle_lookup, seed=4242)


def _load(snapshots, tag):
    return snapshots["x" + tag].astype(np.float64), snapshots["v" + tag].astype(np.float64), snapshots["t" + tag].astype(np.int32)


@pytest.mark.parametrize("tag,pair", PHASES)
@pytest.mark.parametrize("use_arrays", [False, True])
def test_forces_and_energies(eng, orc, oracle_lib, snapshots, tag, pair, use_arrays):
    x, v, t = _load(snapshots, tag)
    orc.set_state(t, x, v)
    orc.set_pair(pair)
    if pair == "soft":
        orc.L.esbo_pair_soft(orc.h, 54.0, oracle_lib._dp(orc.soft_cut))
    fo, eo, st = orc.forces()
    assert st == 0
    eng.upload_state(np.stack([x] * 3), np.stack([v] * 3), np.stack([t] * 3))
    fg, eg = eng.forces(pair, soft_a=54.0, use_arrays=use_arrays)
    scale = np.abs(fo).max()
    assert np.abs(fg[0] - fo).max() / scale < TOL
    assert np.array_equal(fg[0], fg[1]) and np.array_equal(fg[0], fg[2])     # replicas are independent and deterministic
    assert np.allclose(eg[0][:3], eo[:3], rtol=2e-6, atol=1e-3)
    assert np.all(eng.status() == 0)


@pytest.mark.parametrize("tag,pair", PHASES[1:])
def test_post_force_langevin_walls_anchors(eng, orc, snapshots, tag, pair, box11):
    x, v, t = _load(snapshots, tag)
    orc.set_state(t, x, v)
    orc.set_pair(pair)
    orc.set_seed(4242, 77)
    fo, eo, st = orc.forces(with_post=True, with_noise=True)
    eng.upload_state(np.stack([x] * 3), np.stack([v] * 3), np.stack([t] * 3))
    fg, eg = eng.forces(pair, with_post=True, eval_index=77)
    assert np.abs(fg[0] - fo).max() / np.abs(fo).max() < TOL
    assert np.all(fg[:, box11.anchors()] == 0.0)
    assert not np.array_equal(fg[0], fg[1])                                   # different seeds, different noise
    assert abs(eg[0][3] - eo[3]) <= 2e-6 * abs(eo[3]) + 1e-6


def test_bin_index_is_exact(eng, orc, snapshots, box11):
    """The LOOKUP bin of every pair is the oracle's: with array lookups the only difference left is
    fp32 rounding of the table values and of the accumulation (<< one bin step of ~1e-4 relative)."""
    x, v, t = _load(snapshots, "1999")
    orc.set_state(t, x, v)
    orc.set_pair("B")
    fo, _, _ = orc.forces()
    eng.upload_state(np.stack([x] * 3), np.stack([v] * 3), np.stack([t] * 3))
    fg, _ = eng.forces("B", use_arrays=True)
    # one flipped bin on a strongly repulsive pair (F ~ 50, neighbouring bins differ by ~1e-4 relative)
    # would show up as ~4e-5 of the largest force; fp32 table values and the fixed-point accumulation (2^-16 per
    # pair and component, ~60 pairs per bead: ~1e-4 absolute at worst) stay below 3e-6 on the beads without bonded
    # terms (most pairs have such a bead: RBP, RNP, Ser5P); the polymer beads add the rounding of their bonded shares
    # (2^-12 per term and component, <= 5 terms: 6e-4 absolute at worst) and stay below 2e-5
    bonded = np.zeros(len(x), bool)
    bonded[np.unique(box11.bonds[:, 1:])] = True
    err = np.abs(fg[0] - fo).max(axis=1) / np.abs(fo).max()
    assert err[~bonded].max() < 3e-6 and err[bonded].max() < 2e-5, (err[~bonded].max(), err[bonded].max())


def test_pairs_excluded_and_counted(eng, orc, snapshots, box11):
    """1-2/1-3/1-4 partners do not interact: moving a bonded neighbour into overlap changes only bonded forces."""
    x, v, t = _load(snapshots, "500")
    x = x.copy()
    x[61] = x[60] + np.float32(0.05)                 # 1-2 pair almost on top of each other
    x[63] = x[60] + np.float32(0.07)                 # 1-4 pair
    orc.set_state(t, x, v)
    orc.set_pair("B")
    fo, _, _ = orc.forces()
    eng.upload_state(np.stack([x] * 3), np.stack([v] * 3), np.stack([t] * 3))
    fg, _ = eng.forces("B")
    assert np.abs(fg[0] - fo).max() / np.abs(fo).max() < TOL


def test_generic_size_path(oracle_lib, pair_tables, oracle_lookup):
    """A bead count outside the instantiated capacities (box 13: 1516 beads -> n_pad 1536) runs the generic
    instantiation of the kernels: soft-phase forces at the initial condition and a short run against the oracle."""
    from enhancershoebox_b200 import datafile, protocol as P
    d = datafile.make_shoebox(13, 3, False, seed=2)
    eng = E.ShoeboxBatch(d, n_replicas=2, seeds=[1, 2])
    assert eng.n_pad not in (1056, 1184, 1312, 1408, 1600)
    orc = oracle_lib.Oracle(d, pair_tables, oracle_lookup, seed=1)
    orc.set_pair("soft")
    orc.L.esbo_pair_soft(orc.h, 30.0, oracle_lib._dp(orc.soft_cut))
    fo, eo, st = orc.forces()
    fg, eg = eng.forces("soft", soft_a=30.0)
    assert st == 0 and np.abs(fg[0] - fo).max() / np.abs(fo).max() < TOL
    eng.set_schedule(12, observe=True)
    eng.run(0, 12)                                        # soft phase + the first table chunks
    assert np.all(eng.status() == 0) and eng.step == 2400
    x, v, t = eng.state()
    orc.set_state(t[0], x[0], v[0])
    orc.L.esbo_set_bond_coeff(orc.h, 1, P.BOND_K, float(eng.bond_r0()[1]))
    orc.set_pair("A")
    fo, _, st = orc.forces()
    fg, _ = eng.forces("A")
    assert st == 0 and np.abs(fg[0] - fo).max() / np.abs(fo).max() < TOL
    eng.close()
