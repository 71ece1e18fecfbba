"""The CPU oracle itself: known answers, internal consistency, LAMMPS-semantics details."""
import numpy as np
import pytest

from enhancershoebox_b200 import datafile, ljsoft, protocol as P


def test_philox_known_answers(oracle_lib):
    # Random123 kat_vectors, philox4x32-10
    assert oracle_lib.philox((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle_lib.philox((0xffffffff,) * 4, (0xffffffff,) * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle_lib.philox((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_lookup_table_matches_closed_form_away_from_rc(oracle_lib, pair_tables, sections, oracle_lookup):
    for (key, cut), (e, f, inner, delta) in zip(pair_tables.specs, oracle_lookup):
        g = ljsoft.section_geometry(key)
        assert inner == 1e-8 and abs(delta - (cut * cut - 1e-8) / 29999) < 1e-18 and len(f) == 29999
        mid = inner + (np.arange(len(f)) + 0.5) * delta
        r = np.sqrt(mid)
        x = g.c + (r / g.s) ** 6
        fc = np.where(r < g.rc, 24 * g.eps_eff * r ** 4 / g.s ** 6 * (2 / x ** 3 - 1 / x ** 2), 0.0)
        irc = int((min(g.rc, cut) ** 2 - inner) / delta)
        lo, hi = 2500, max(irc - 1500, 2501)
        assert np.max(np.abs(f[lo:hi] - fc[lo:hi]) / (1 + np.abs(fc[lo:hi]))) < 1e-8, key   # spline == closed form
        if irc + 1500 < len(f):
            assert np.max(np.abs(f[irc + 1500:])) < 1e-9, key                                  # zero beyond rc
    with pytest.raises(ValueError):
        oracle_lib.build_lookup_table(sections["RBP_RBP"], 3.5)                              # cut > rhi: LAMMPS errors out


def test_special_lists(oracle_lib, box11, pair_tables, oracle_lookup):
    o = oracle_lib.Oracle(box11, pair_tables, oracle_lookup)
    start, sp = o.special()
    n_left = 123
    assert sorted(sp[start[0]:start[1]]) == [1, 2, 3]                      # chain end
    assert sorted(sp[start[60]:start[61]]) == [57, 58, 59, 61, 62, 63]     # interior bead: 1-2, 1-3, 1-4
    assert sorted(sp[start[n_left - 1]:start[n_left]]) == [n_left - 4, n_left - 3, n_left - 2]   # no bond across the two polymers
    assert start[370] == start[-1]                                          # free species have no exclusions


def _energy(o):
    f, e, st = o.forces()
    return e.sum(), f


@pytest.mark.parametrize("pair", ["soft", None])
def test_forces_are_minus_energy_gradient(oracle_lib, box11, pair_tables, oracle_lookup, snapshots, pair):
    """pair soft / bonds / angles: F = -dE/dx by central differences (walls are tested separately)."""
    o = oracle_lib.Oracle(box11, pair_tables, oracle_lookup)
    x = snapshots["x9"].astype(np.float64)
    t = snapshots["t9"].astype(np.int32)
    o.set_state(t, x)
    if pair == "soft":
        o.L.esbo_pair_soft(o.h, 30.0, oracle_lib._dp(o.soft_cut))
    e0, f = _energy(o)
    rng = np.random.default_rng(0)
    h = 1e-6
    for i in list(rng.integers(0, 370, 6)) + list(rng.integers(370, 1290, 4)):
        for d in range(3):
            xp, xm = x.copy(), x.copy()
            xp[i, d] += h
            xm[i, d] -= h
            o.set_state(t, xp)
            ep, _ = _energy(o)
            o.set_state(t, xm)
            em, _ = _energy(o)
            assert abs(-(ep - em) / (2 * h) - f[i, d]) < 2e-4 * (1 + abs(f[i, d])), (i, d)


def test_wall_and_anchor_post_force(oracle_lib, box11, pair_tables, oracle_lookup):
    o = oracle_lib.Oracle(box11, pair_tables, oracle_lookup)
    o.set_langevin(0)
    x = box11.x.copy()
    x[400] = [12.5, 0.0, 0.0]          # 0.5 from xhi = 13  -> f_x = -2*100*(3-0.5)
    x[401] = [0.0, -9.0, 6.9]          # 0.5 from ylo, 0.1 from zhi
    o.set_state(box11.types, x)
    f0, _, _ = o.forces(with_post=False)
    f1, e1, st = o.forces(with_post=True)
    assert st == 0
    assert abs((f1 - f0)[400, 0] + 500.0) < 1e-9 and abs((f1 - f0)[401, 1] - 500.0) < 1e-9 and abs((f1 - f0)[401, 2] + 580.0) < 1e-9
    assert np.all(f1[box11.anchors()] == 0.0)                      # fix setforce 0 0 0 is applied last
    assert e1[3] > 0
    x[402] = [13.0, 0, 0]                                           # on the wall: LAMMPS errors out
    o.set_state(box11.types, x)
    assert o.forces(with_post=True)[2] == 1


def test_adapt_ramps_and_reset_no(oracle_lib, box11, pair_tables, oracle_lookup):
    """fix adapt: A = 60 n/2000 every step, r0 = 0.033 + 0.967 n/2000 every 10 steps; values stick after unfix."""
    o = oracle_lib.Oracle(box11, pair_tables, oracle_lookup, seed=3)
    plans = P.chunk_schedule("genes_stages", n_chunks=11)
    assert o.run_chunk(plans[0]) == 0
    assert o.step == 200 and abs(o.soft_a() - 6.0) < 1e-12 and abs(o.bond_r0(1) - (0.033 + 0.967 * 0.1)) < 1e-12
    for pl in plans[1:9]:
        assert o.run_chunk(pl) == 0
    assert o.step == 1800 and abs(o.bond_r0(1) - 0.9033) < 1e-12 and abs(o.soft_a() - 54.0) < 1e-12
    assert o.run_chunk(plans[9]) == 0 and o.run_chunk(plans[10]) == 0
    assert abs(o.bond_r0(1) - 0.9033) < 1e-12 and o.bond_r0(2) == 0.3        # reset no: r0 stays at 0.9033, not 1.0
    x, v, f, t = o.state()
    assert np.all(x[box11.anchors()] == box11.x[box11.anchors()])            # anchors never move
    ke = 0.5 * (v ** 2).sum() / box11.n_atoms
    assert 1.3 < ke < 1.7                                                    # 3/2 kT with T = 1


def test_shoebox_driver_actin_r0_re_ramps(oracle_lib, pair_tables, oracle_lookup):
    """fix s3 is never unfixed in run_single_shoebox.py: actin bond r0 ramps 0.033 -> 0.3 in every later run."""
    d = datafile.make_shoebox(9, 3, False, seed=1)
    o = oracle_lib.Oracle(d, pair_tables, oracle_lookup, seed=3)
    plans = P.chunk_schedule("shoebox", n_chunks=11)
    o.apply_plan(plans[10])
    o.L.esbo_run(o.h, 100, -1, -1)
    assert abs(o.bond_r0(2) - (0.033 + (0.3 - 0.033) * 100 / 100)) < 1e-12   # a 100-step run ramps over its own length
    o.apply_plan(plans[10])
    o.L.esbo_run(o.h, 20, -1, -1)
    assert abs(o.bond_r0(2) - 0.3) < 1e-12


def test_noise_is_reproducible_and_keyed(oracle_lib, box11, pair_tables, oracle_lookup):
    o = oracle_lib.Oracle(box11, pair_tables, oracle_lookup, seed=11)
    o.set_pair("soft")
    o.set_seed(11, 5)
    f1, _, _ = o.forces(with_post=True, with_noise=True)
    o.set_seed(11, 5)
    f2, _, _ = o.forces(with_post=True, with_noise=True)
    o.set_seed(11, 6)
    f3, _, _ = o.forces(with_post=True, with_noise=True)
    o.set_seed(11, 5)
    f0, _, _ = o.forces(with_post=True, with_noise=False)
    assert np.array_equal(f1, f2) and not np.array_equal(f1, f3)
    noise = (f1 - f0)[~np.isin(np.arange(box11.n_atoms), box11.anchors())] / np.sqrt(9600)
    assert np.all(np.abs(noise) <= 0.5 + 1e-9) and abs(noise.mean()) < 0.02 and abs(noise.var() - 1 / 12) < 0.01


def test_neighbour_delay_semantics(oracle_lib, box11, pair_tables, oracle_lookup, snapshots):
    """neigh_modify every 1 delay 10 check yes (run_single_shoebox.py:533): every run builds at set-up, afterwards
    never sooner than 10 steps after the last build; with delay 0 the list follows the skin/2 rule alone.  The
    extra builds of delay 0 show that most delay-10 builds were already due ("dangerous builds")."""
    x, v, t = snapshots["x500"].astype(np.float64), snapshots["v500"].astype(np.float64), snapshots["t500"].astype(np.int32)
    plan = P.chunk_schedule("genes_stages", n_chunks=501)[500]
    counts = {}
    for delay in (10, 0):
        o = oracle_lib.Oracle(box11, pair_tables, oracle_lookup, seed=4, neigh_delay=delay)
        o.set_state(t, x, v)
        o.apply_plan(plan)
        o.reset_counters()
        assert o.run(200) == 0
        counts[delay] = o.counters()
    c10, c0 = counts[10], counts[0]
    assert c10["evals"] == c0["evals"] == 201
    assert 2 <= c10["builds"] <= 1 + 200 // 10                     # set-up + at most one per 10 steps
    assert c0["builds"] > c10["builds"] and c0["dangerous"] == 0
    assert 0 < c10["dangerous"] <= c10["builds"] - 1


def test_lookup_table_against_scipy_cubic_spline(oracle_lib, sections):
    """The table construction ([LAMMPS-ext] PairTable::read_table with RSQ, spline_table, compute_table) restated with a
    third-party spline: scipy's CubicSpline with clamped ends is the same interpolant as LAMMPS' spline()/splint()
    (Numerical Recipes).  End slopes: E' = -F at both ends, F' by finite differences of the first / last two file rows."""
    from scipy.interpolate import CubicSpline
    for key, cut in (("SER5P_SER5P", 0.9), ("RNP_RNP_ATTRACTION", 2.9), ("CHROMATIN_CHROMATIN", 1.2)):
        sec = sections[key]
        n, rlo, rhi = int(sec.n), float(sec.rlo), float(sec.rhi)
        ef, ff = np.asarray(sec.e, np.float64), np.asarray(sec.f, np.float64)
        r = np.sqrt(rlo * rlo + (rhi * rhi - rlo * rlo) * np.arange(n) / (n - 1))       # RSQ: file r values are regenerated
        e_spl = CubicSpline(r, ef, bc_type=((1, -ff[0]), (1, -ff[-1])))
        f_spl = CubicSpline(r, ff, bc_type=((1, (ff[1] - ff[0]) / (r[1] - r[0])), (1, (ff[-1] - ff[-2]) / (r[-1] - r[-2]))))
        e, f, inner, delta = oracle_lib.build_lookup_table(sec, cut)
        assert inner == rlo * rlo and abs(delta - (cut * cut - inner) / 29999) < 1e-18
        rm = np.sqrt(inner + (np.arange(29999) + 0.5) * delta)                          # bin midpoints in r
        e_ref, f_ref = e_spl(rm), f_spl(rm) / rm
        assert np.max(np.abs(e - e_ref) / (1 + np.abs(e_ref))) < 1e-9, key
        assert np.max(np.abs(f - f_ref) / (1 + np.abs(f_ref))) < 1e-9, key


def test_nve_conserves_energy_without_thermostat(oracle_lib, box11, pair_tables, oracle_lookup, snapshots):
    """fix nve alone (langevin off, constant soft prefactor, no fix adapt): velocity-Verlet with forces that are the
    gradient of the reported energies (pair soft + bonds + angles + walls) conserves E_pot + E_kin to O(dt^2) --
    an integrator/force-consistency check that does not depend on LAMMPS."""
    x, v, t = snapshots["x9"].astype(np.float64), snapshots["v9"].astype(np.float64), snapshots["t9"].astype(np.int32)
    o = oracle_lib.Oracle(box11, pair_tables, oracle_lookup, seed=3)
    o.set_langevin(0)
    o.set_state(t, x, v)
    o.set_pair("soft")
    o.L.esbo_pair_soft(o.h, 54.0, oracle_lib._dp(o.soft_cut))
    o.L.esbo_clear_adapt(o.h)

    def energies():
        _, e, st = o.forces(with_post=True)
        _, vv, _, _ = o.state()
        return st, e.sum() + 0.5 * (vv ** 2).sum(), 0.5 * (vv ** 2).sum()

    st, e0, ke0 = energies()
    assert st == 0 and ke0 > 1000
    for _ in range(4):
        assert o.run(50) == 0
        st, e1, _ = energies()
        assert st == 0 and abs(e1 - e0) < 2e-3 * ke0, (e0, e1)
