"""configs[1]: box-11 actin shoebox (1593 beads, run_single_shoebox.py command stream) with static topology --
the reference's own `LatB` condition switches polymerisation off (run_single_shoebox.py:459-460).
Covers the actin pair_coeff rows (appendix C.2), bond/angle type 2 and the never-unfixed fix s3 that
re-ramps the actin bond r0 inside every later run (:638-639 vs :794-795)."""
import numpy as np
import pytest

from enhancershoebox_b200 import datafile, engine as E, ljsoft, protocol as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def actin_setup(oracle_lib):
    d = datafile.make_shoebox(11, 3, True, seed=11)
    pt = P.pair_tables("shoebox", "LatB")
    secs = {k: ljsoft.generated_section(k) for k in sorted({s[0] for s in pt.specs})}
    lookup = [oracle_lib.build_lookup_table(secs[k], c) for (k, c) in pt.specs]
    return d, pt, lookup


def test_actin_batch_runs_and_matches_oracle(oracle_lib, actin_setup):
    d, pt, lookup = actin_setup
    assert d.n_atoms == 1593
    R = 4
    eng = E.ShoeboxBatch(d, n_replicas=R, driver="shoebox", condition="LatB", seeds=[3, 4, 5, 3])
    assert eng.n_pad == 1600
    eng.set_schedule(22, observe=True)
    eng.run(0, 20)                                   # soft phase, table phase A, Ser5P switch-on
    assert np.all(eng.status() == 0)
    assert abs(eng.bond_r0()[1] - 0.9033) < 1e-12 and abs(eng.bond_r0()[2] - 0.3) < 1e-12    # s3 ends every run at 0.3
    x, v, t = eng.state()
    assert np.array_equal(x[0], x[3]) and not np.array_equal(x[0], x[1])
    ke = 0.5 * (v ** 2).sum(axis=(1, 2)) / d.n_atoms
    assert np.all((ke > 1.3) & (ke < 1.7))
    # force parity at the equilibrated actin state (table phase B of the shoebox driver)
    orc = oracle_lib.Oracle(d, pt, lookup, seed=3)
    orc.set_state(t[0], x[0], v[0])
    orc.L.esbo_set_bond_coeff(orc.h, 1, P.BOND_K, float(eng.bond_r0()[1]))      # fix adapt left r0 = 0.9033 behind
    orc.set_pair("B")
    fo, eo, st = orc.forces()
    fg, eg = eng.forces("B")
    assert st == 0 and np.abs(fg[0] - fo).max() / np.abs(fo).max() < 1e-5
    assert np.allclose(eg[0][:3], eo[:3], rtol=2e-6, atol=1e-3)
    # one more chunk of 25 steps on both sides, shared Philox noise, fix s3 re-ramping 0.033 -> 0.3 over the run
    plan = P.chunk_schedule("shoebox", "LatB", n_chunks=21)[20]
    plan = P.ChunkPlan(plan.index, plan.pair, plan.ramp, plan.adapt_soft, plan.adapt_bond1, plan.adapt_bond2, 25)
    assert plan.adapt_bond2 and plan.pair == "B"
    eng.set_schedule(plans=[plan], observe=False)
    eng.run(0, 1)
    xg, vg, _ = eng.state()
    orc.set_seed(3, int(20 * 201))                   # 20 chunks x (200 steps + set-up) evaluations so far
    orc.apply_plan(plan)
    assert orc.run(25) == 0
    xo, vo, _, _ = orc.state()
    dx = np.abs(xg[0] - xo).max(axis=1)
    assert np.median(dx) < 1e-5 and np.mean(dx < 1e-4) > 0.99 and dx.max() < 0.05
    assert abs(eng.bond_r0()[2] - orc.bond_r0(2)) < 1e-12 and abs(orc.bond_r0(2) - (0.033 + 0.267 * 20 / 25)) < 1e-12
    eng.close()


def test_polymerising_conditions_are_refused():
    from enhancershoebox_b200 import drivers
    a = drivers.parse_args("-b 11 -r 1 -o out/ -c Control -p 3 -a 30 -x 80 -z 300 -n 0.001".split(), with_actin=True)
    with pytest.raises(SystemExit, match="polymerisation"):
        drivers.run_condition("shoebox", a, n_chunks=1)
