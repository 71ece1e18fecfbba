"""(f)4: treatments of run_single_shoebox.py:876-887 -- at i == NRuns Hexanediol / JQ-1 swap pair tables (map "T"),
Flavopiridol stops gene activation; the run is extended by treatment_duration chunks (:379-385)."""
import numpy as np
import pytest

from enhancershoebox_b200 import engine as E, ljsoft, protocol as P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("condition", ["Hexanediol", "JQ-1"])
def test_treatment_tables_and_schedule(oracle_lib, box11, snapshots, condition):
    x, v, t = snapshots["x1999"].astype(np.float64), snapshots["v1999"].astype(np.float64), snapshots["t1999"].astype(np.int32)
    pt = P.pair_tables("shoebox", condition, ("A", "B", "T"))
    secs = {k: ljsoft.generated_section(k) for k in sorted({s[0] for s in pt.specs})}
    lookup = [oracle_lib.build_lookup_table(secs[k], c) for (k, c) in pt.specs]
    eng = E.ShoeboxBatch(box11, n_replicas=2, driver="shoebox", condition=condition, seeds=[8, 9])
    eng.upload_state(np.stack([x, x]), np.stack([v, v]), np.stack([t, t]))
    orc = oracle_lib.Oracle(box11, pt, lookup, seed=8)
    orc.set_state(t, x, v)
    f = {}
    for pair in ("B", "T"):                                   # force parity before and after the treatment's pair_coeff commands
        orc.set_pair(pair)
        fo, eo, st = orc.forces()
        fg, eg = eng.forces(pair)
        assert st == 0 and np.abs(fg[0] - fo).max() / np.abs(fo).max() < 1e-5
        f[pair] = fo
    s5p = np.flatnonzero(t == 8)
    assert np.abs(f["T"][s5p] - f["B"][s5p]).max() > 1.0        # the Ser5P interactions did change
    # the schedule switches to map T at i == NRuns and runs treatment_duration more chunks
    plans = P.chunk_schedule("shoebox", condition)
    assert len(plans) == P.N_RUNS + P.TREATMENT_DURATION[condition]
    assert plans[P.N_RUNS - 1].pair == "B" and plans[P.N_RUNS].pair == "T"
    eng.set_schedule(plans=plans[P.N_RUNS - 2:P.N_RUNS + 3], observe=True)
    eng.run()
    assert np.all(eng.status() == 0)
    fr = eng.frames()
    assert np.array_equal(fr[0, :, 0], (np.arange(P.N_RUNS - 2, P.N_RUNS + 3) + 1) * 6.0)
    eng.close()


def test_flavopiridol_stops_activation():
    pl = P.chunk_schedule("shoebox", "Flavopiridol")
    assert len(pl) == P.N_RUNS + P.TREATMENT_DURATION["Flavopiridol"] and all(p.pair == "B" for p in pl[15:])
