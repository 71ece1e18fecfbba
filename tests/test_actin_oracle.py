"""oracle/actin.py (restatement of run_single_shoebox.py:969-1298): bookkeeping invariants and the event schedule.
The device is compared with it event by event in tests/test_gpu_actin.py."""
import numpy as np

from enhancershoebox_b200 import datafile, protocol as P
from oracle import actin as A


def test_event_schedule_flags():
    assert [P.actin_flags(i) for i in (0, 38, 39, 40, 41, 44, 45, 79, 80)] == [0, 0, 1, 7, 1, 1, 5, 1, 7]
    assert P.actin_flags(500, "LatB") == 0 and P.actin_flags(500, "SwinA_nopol") == 0      # t_polymerization_on = 2*NRuns


def test_bookkeeping_invariants_over_many_events():
    d = datafile.make_shoebox(11, 3, True, seed=3)
    ids = np.flatnonzero(d.types == P.ACTIN_TYPE)
    first, count = int(ids[0]), len(ids)
    assert count == 303
    st = A.ActinState(list(range(first + 3, first + count)), [[first, first + 1, first + 2]], seed=5, p_nucleation=0.05)
    assert st.bonds == {(first, first + 1), (first + 1, first + 2)} and st.angles == {(first, first + 1, first + 2)}
    x = d.x.astype(np.float32)
    rng = np.random.default_rng(1)
    seen_nucleation = seen_growth = seen_release = False
    for i in range(39, 39 + 30):
        n_fil, n_in = len(st.filament_details), sum(len(f) for f in st.filament_details)
        x0 = x.copy()
        calls = st.events(i, x, P.actin_flags(i))
        assert calls >= len(st.free_actin)                                   # at least one draw per free monomer (nucleation loop)
        lens = [len(f) for f in st.filament_details]
        assert all(v >= 2 for v in lens) and len(st.free_actin) + sum(lens) == count
        members = st.free_actin + [a for f in st.filament_details for a in f]
        assert sorted(members) == list(range(first, first + count))           # every actin bead is free or in exactly one filament
        assert len(st.bonds) == sum(v - 1 for v in lens) and len(st.angles) == sum(max(v - 2, 0) for v in lens)
        for f in st.filament_details:                                          # exclusions follow the chain to depth 3
            assert st.exclusions(f[0]) == set(f[1:4])
        moved = np.flatnonzero(np.any(x != x0, axis=1))
        assert np.all((moved >= first) & (moved < first + count))             # only actin monomers are relocated
        seen_nucleation |= len(st.filament_details) > n_fil
        seen_growth |= sum(lens) > n_in
        seen_release |= bool(P.actin_flags(i) & 4) and len(lens) > 0
        x[first:first + count] += rng.normal(0, 0.05, (count, 3)).astype(np.float32)     # some motion between events
    assert seen_nucleation and seen_growth and seen_release
    assert st.stats()[0] == len(st.filament_details) and st.stats()[3] == len(st.free_actin)


def test_same_seed_same_events():
    d = datafile.make_shoebox(11, 3, True, seed=3)
    first = int(np.flatnonzero(d.types == P.ACTIN_TYPE)[0])
    out = []
    for seed in (9, 9, 10):
        st = A.ActinState(list(range(first + 3, first + 303)), [[first, first + 1, first + 2]], seed=seed, p_nucleation=0.1)
        x = d.x.astype(np.float32)
        for i in (39, 40):
            st.events(i, x, P.actin_flags(i))
        out.append((st.free_actin, st.filament_details, x))
    assert out[0][0] == out[1][0] and out[0][1] == out[1][1] and np.array_equal(out[0][2], out[1][2])
    assert out[0][1] != out[2][1]
