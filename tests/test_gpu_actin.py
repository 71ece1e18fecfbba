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


def _event_plans(indices):
    """Zero-step chunks that only run what the drivers do between two `run` commands, for the given Python timesteps."""
    out = []
    for i in indices:
        pl = P.chunk_schedule("shoebox", "Control", n_chunks=i + 1)[i]
        out.append(P.ChunkPlan(pl.index, pl.pair, None, False, False, False, 0))      # (no fix adapt: a 0-step ramp is 0/0)
    return out


def test_actin_events_match_the_restatement(oracle_lib, actin_setup):
    """a20: nucleation, plus/minus-end polymerisation and minus-/plus-end depolymerisation on the device against
    oracle/actin.py (run_single_shoebox.py:969-1298 with the shared Philox stream), event by event on the same
    coordinates: relocated positions bit for bit, free_actin / filament_details in the reference's list order;
    then the topology the device derived (bonds, angles, 1-2/1-3/1-4 exclusions) through the force parity hook."""
    from oracle import actin as A
    d, _, _ = actin_setup
    pt = P.pair_tables("shoebox", "Control")
    secs = {k: ljsoft.generated_section(k) for k in sorted({s[0] for s in pt.specs})}
    lookup = [oracle_lib.build_lookup_table(secs[k], c) for (k, c) in pt.specs]
    R, p_nuc = 3, 0.05
    seeds = [11, 12, 11]
    eng = E.ShoeboxBatch(d, n_replicas=R, driver="shoebox", condition="Control", seeds=seeds, p_nucleation=p_nuc)
    first, count = eng.actin.first, eng.actin.count
    assert count == 303 and np.all(d.types[first:first + count] == 3)
    eng.set_schedule(22, observe=True)
    eng.run(0, 22)                                    # soft phase ... tables B: (i+1) < 40, no events yet
    st0 = eng.actin_state()
    assert st0["filaments"][0] == [[first, first + 1, first + 2]] and st0["free"][0] == list(range(first + 3, first + count))
    stats, _ = eng.actin_frames(0, 22)
    assert np.all(stats[:, :, 0] == 1) and np.all(stats[:, :, 1] == 3) and np.all(stats[:, :, 3] == 300)
    orc = [A.ActinState(st0["free"][r], st0["filaments"][r], seeds[r], p_nuc) for r in range(R)]
    # alternate event-only chunks (Python timesteps 39..46 and 80: flags 1, 7, 1, ..., 5 and 7) with 40 MD steps
    md = P.chunk_schedule("shoebox", "LatB", n_chunks=21)[20]
    md = P.ChunkPlan(md.index, md.pair, md.ramp, md.adapt_soft, md.adapt_bond1, md.adapt_bond2, 40)
    n_nucleated = n_grown = n_released = 0
    assert P.actin_flags(4000) == 15 and P.actin_flags(3999) == 1        # plus-end release: i % t_off_plus == 0 only (:496, 1220)
    for i in (39, 40, 41, 42, 43, 44, 45, 46, 80, 4000):
        x_pre, _, _ = eng.state()
        eng.set_schedule(plans=_event_plans([i]), observe=True)
        eng.run(0, 1)
        x_post, _, _ = eng.state()
        st = eng.actin_state()
        stats, hist = eng.actin_frames(0, 1)
        for r in range(R):
            before = (len(orc[r].filament_details), sum(len(f) for f in orc[r].filament_details))
            xo = x_pre[r].astype(np.float32).copy()
            orc[r].events(i, xo, P.actin_flags(i))
            assert st["free"][r] == orc[r].free_actin, (i, r)
            assert st["filaments"][r] == orc[r].filament_details, (i, r)
            assert np.array_equal(x_post[r].astype(np.float32), xo), (i, r)
            assert list(stats[r, 0]) == orc[r].stats(), (i, r)
            assert hist[r, 0].sum() <= count
            after = (len(orc[r].filament_details), sum(len(f) for f in orc[r].filament_details))
            n_nucleated += max(after[0] - before[0], 0); n_grown += max(after[1] - before[1], 0); n_released += P.actin_flags(i) & 4 and 1
            if i == 4000:                                   # both ends released: every surviving filament lost two beads (plus what it grew)
                assert before[0] > 0 and all(len(f) >= 2 for f in orc[r].filament_details)
        assert st["free"][0] == st["free"][2] and st["free"][0] != st["free"][1]          # same seed, same events
        eng.set_schedule(plans=[md], observe=False)                                    # move on: 40 MD steps, no events
        eng.run(0, 1)
        assert np.all(eng.status() == 0), (i, eng.status())
    assert n_nucleated > 10 and n_grown > 30 and n_released > 0
    # the topology behind the next run: forces of the device against the oracle built from the restatement's bonds/angles
    x, v, t = eng.state()
    import dataclasses
    for r in (0, 1):
        bonds = [b for b in d.bonds if b[1] < first] + [[2, a, b] for (a, b) in sorted(orc[r].bonds)]
        angles = [g for g in d.angles if g[1] < first] + [[2, a, b, c] for (a, b, c) in sorted(orc[r].angles)]
        dr = dataclasses.replace(d, bonds=np.array(bonds, np.int32), angles=np.array(angles, np.int32))
        o = oracle_lib.Oracle(dr, pt, lookup, seed=seeds[r])
        o.set_state(t[r], x[r], v[r])
        o.L.esbo_set_bond_coeff(o.h, 1, P.BOND_K, float(eng.bond_r0()[1]))
        o.L.esbo_set_bond_coeff(o.h, 2, P.BOND_K, float(eng.bond_r0()[2]))
        o.set_pair("B")
        fo, eo, sto = o.forces()
        fg, eg = eng.forces("B")
        assert sto == 0 and np.abs(fg[r] - fo).max() / np.abs(fo).max() < 1e-5, r
        assert np.allclose(eg[r][:3], eo[:3], rtol=2e-6, atol=1e-3)
    eng.close()


def test_actin_driver_writes_actin_files(tmp_path, monkeypatch):
    """The shoebox driver with polymerisation on: actinTrack.txt / actinAroundCluster.txt next to geneTrack.txt."""
    from enhancershoebox_b200 import drivers
    monkeypatch.chdir(tmp_path)
    a = drivers.parse_args(f"-b 11 -r 1 -o {tmp_path}/out/ -c Control -p 3 -a 30 -x 80 -z 300 -n 0.05".split(), with_actin=True)
    d = datafile.make_shoebox(11, 3, True, seed=11)
    assert drivers.run_condition("shoebox", a, n_chunks=46, data=d, quiet=True) == 0
    run = tmp_path / "out" / "run1"
    rows = (run / "actinTrack.txt").read_text().splitlines()
    assert len(rows) == 46 and rows[0].split("\t") == ["6.0", "1", "3.0", "0.0", "300"]
    last = rows[-1].split("\t")
    assert int(last[1]) >= 1 and int(last[4]) < 300                      # filaments formed, monomers consumed
    assert len((run / "actinAroundCluster.txt").read_text().splitlines()) == 47
    assert len((run / "geneTrack.txt").read_text().splitlines()) == 46
