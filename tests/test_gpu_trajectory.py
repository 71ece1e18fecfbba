"""Level 2: short trajectories of the step kernel against the oracle, chunk semantics, failure flags."""
import numpy as np
import pytest

from enhancershoebox_b200 import engine as E
from enhancershoebox_b200 import protocol as P

pytestmark = pytest.mark.gpu


def _plans(first, count, n_steps):
    pl = P.chunk_schedule("genes_stages", n_chunks=first + count)[first:]
    return [P.ChunkPlan(p.index, p.pair, p.ramp, p.adapt_soft, p.adapt_bond1, p.adapt_bond2, n_steps) for p in pl]


def _state(snapshots, tag):
    return snapshots["x" + tag].astype(np.float64), snapshots["v" + tag].astype(np.float64), snapshots["t" + tag].astype(np.int32)


@pytest.mark.parametrize("n_steps", [25, 100])
@pytest.mark.parametrize("mode", [2, 1])
def test_short_trajectory_matches_oracle(oracle_lib, box11, pair_tables, oracle_lookup, snapshots, mode, n_steps):
    """25 and 100 steps from an equilibrated frame, zero-noise (mode 2: drag only) and with the shared
    Philox noise (mode 1), fp32 state on the device vs the fp64 oracle.

    Stated tolerance (absolute, bead diameter = 1): after 25 steps the median |dx| is < 1e-5 and 99 % of
    the beads are within 1e-4; after 100 steps the median is still < 1e-5 and at least 60 % are within
    1e-4.  A tight bound on every bead is not meaningful: LOOKUP tables are piecewise constant and the
    potential is CUT, not shifted (the force jumps by up to ~1.6 at rc, ljsoft_potential.py:11-15), so a
    1e-7 roundoff difference moves a pair across a bin edge / rc one step earlier or later and the
    kick spreads through the dense Ser5P cluster.  The oracle itself, restarted from coordinates
    perturbed by 1e-7, shows the same figures (25 steps: 100 % within 1e-4, max 4e-5; 100 steps:
    80 % within 1e-4, max 1.2e-3), i.e. two LAMMPS builds would differ just as much."""
    x, v, t = _state(snapshots, "500")
    eng = E.ShoeboxBatch(box11, n_replicas=2, seeds=[99, 99], langevin_mode=mode)
    eng.upload_state(np.stack([x, x]), np.stack([v, v]), np.stack([t, t]))
    plans = _plans(500, 1, n_steps)
    eng.set_schedule(plans=plans, observe=False)
    eng.run(0, 1)
    xg, vg, tg = eng.state()
    orc = oracle_lib.Oracle(box11, pair_tables, oracle_lookup, seed=99)
    orc.set_langevin(mode)
    orc.set_state(t, x, v)
    orc.apply_plan(plans[0])
    assert orc.run(n_steps) == 0
    xo, vo, fo, to = orc.state()
    assert np.all(eng.status() == 0)
    dx = np.abs(xg[0] - xo).max(axis=1)
    frac = np.mean(dx < 1e-4)
    assert np.median(dx) < 1e-5 and dx.max() < 0.05, (np.median(dx), frac, dx.max())
    assert frac > (0.99 if n_steps == 25 else 0.60), (np.median(dx), frac, dx.max())
    assert np.median(np.abs(vg[0] - vo).max(axis=1)) < 1e-4
    assert np.array_equal(xg[0], xg[1]) and np.array_equal(vg[0], vg[1])        # same seed -> same bits
    assert np.abs(xo - x).max() > 0.02                                          # the beads did move
    c = eng.counters()
    assert c["steps"][0] == n_steps and c["evals"][0] == n_steps + 1           # set-up evaluation + one per step
    # neigh_modify every 1 delay 10 check yes: same build schedule as the oracle (a rounding-level difference in a
    # displacement can shift one build by a step), never sooner than 10 steps apart
    oc = orc.counters()
    assert abs(int(c["builds"][0]) - oc["builds"]) <= 1 and c["builds"][0] <= 1 + n_steps // 10
    assert abs(int(c["dangerous"][0]) - oc["dangerous"]) <= 1
    eng.close()


def test_runs_are_bitwise_reproducible(box11, snapshots):
    x, v, t = _state(snapshots, "20")
    out = []
    for _ in range(2):
        eng = E.ShoeboxBatch(box11, n_replicas=4, seeds=[1, 2, 3, 1])
        eng.upload_state(x[None], v[None], t[None], replicate=True)
        eng.set_schedule(plans=_plans(20, 2, 200), observe=True)
        eng.run()
        out.append(eng.state() + (eng.frames(),))
        eng.close()
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)
    xs = out[0][0]
    assert np.array_equal(xs[0], xs[3]) and not np.array_equal(xs[0], xs[1])


def test_soft_phase_ramps_and_step_count(oracle_lib, box11, pair_tables, oracle_lookup):
    """Nine soft chunks from the (violent) initial condition: fix adapt values and step count as in the
    oracle; chromatin r0 freezes at 0.9033 (fix adapt `reset no`)."""
    eng = E.ShoeboxBatch(box11, n_replicas=2, seeds=[5, 6])
    eng.set_schedule(11, observe=False)
    eng.run(0, 1)
    assert eng.step == 200 and abs(eng.bond_r0()[1] - (0.033 + 0.967 * 0.1)) < 1e-12
    eng.run(1, 10)
    assert eng.step == 2200 and abs(eng.bond_r0()[1] - 0.9033) < 1e-12 and eng.bond_r0()[2] == 0.3
    assert np.all(eng.status() == 0)
    x, v, t = eng.state()
    assert np.all(x[:, box11.anchors()] == box11.x[box11.anchors()].astype(np.float32))   # anchors never move
    ke = 0.5 * (v ** 2).sum(axis=(1, 2)) / box11.n_atoms
    assert np.all((ke > 1.3) & (ke < 1.7))                                                # 3/2 kT
    assert np.all(np.abs(x[..., 0]) < 13) and np.all(np.abs(x[..., 1]) < 9.5) and np.all(np.abs(x[..., 2]) < 7)
    eng.close()


def test_wall_crossing_flags_the_replica_only(box11, snapshots):
    x, v, t = _state(snapshots, "20")
    xs = np.stack([x, x, x])
    xs[1, 700, 0] = 13.5                                   # outside xhi: LAMMPS would abort this run
    eng = E.ShoeboxBatch(box11, n_replicas=3, seeds=[1, 1, 1])
    eng.upload_state(xs, np.stack([v] * 3), np.stack([t] * 3))
    eng.set_schedule(plans=_plans(20, 2, 50), observe=True)
    eng.run()
    st = eng.status()
    assert st[0] == 0 and st[2] == 0 and (st[1] & E.ST_WALL)
    xg, _, _ = eng.state()
    assert np.array_equal(xg[0], xg[2]) and not np.array_equal(xg[0], x.astype(np.float32))
    eng.run(1, 1)                                          # a failed replica no longer advances
    assert eng.status()[1] & E.ST_WALL
    eng.close()
