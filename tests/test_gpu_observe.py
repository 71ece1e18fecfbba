"""a14-a18 on the device: the fused per-chunk measurements, gene state machine, RBP->RNP ramp and
frame record are bit-identical to the numpy/scipy restatement on the same coordinates."""
import numpy as np
import pytest

from enhancershoebox_b200 import engine as E
from enhancershoebox_b200 import protocol as P
from oracle import observe as OB

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("promoter", [3, 2, 1])
def test_frames_bit_exact_chunk_by_chunk(snapshots, promoter):
    """Promoter lengths 1 and 2 exercise the probe-index quirk of run_single_shoebox.py:941 (the "promoter probe" is
    atom g - p + 1: the last promoter bead for p = 3, the gene start bead for p = 2, the second gene bead for p = 1)
    and the p-bead promoter mean of d_rp (:946) on the device."""
    from enhancershoebox_b200 import datafile
    box11 = datafile.make_shoebox(11, promoter, False, seed=7)
    R = 4
    x0 = snapshots["x500"].astype(np.float64)                      # (an equilibrated promoter-3 frame: same bead count and chain layout)
    v0 = snapshots["v500"].astype(np.float64)
    t0 = np.asarray(box11.types, np.int32).copy()                  # all RBP, gene inactive, promoter of this length
    seeds = [11, 12, 13, 14]
    thr, act = [5, 5, 1000, 5], [1000, 300, 1000, 0]              # always / sometimes / never (count) / never (p = 0)
    eng = E.ShoeboxBatch(box11, n_replicas=R, seeds=seeds, thresholds=thr, activations=act)
    eng.upload_state(x0[None], v0[None], t0[None], replicate=True)
    first, count = 146, 70
    plans = P.chunk_schedule("genes_stages", n_chunks=first + count)[first:]
    plans = [P.ChunkPlan(p.index, p.pair, p.ramp, p.adapt_soft, p.adapt_bond1, p.adapt_bond2, 20) for p in plans]
    eng.set_schedule(plans=plans, observe=True)
    lay = OB.Layout.from_data(box11)
    exp_rnp = OB.expected_rnp_table(first + count, lay.n_rbprnp, 2000)
    machines = [OB.GeneMachine(lay, threshold=thr[r], p_activation=act[r] / 1000, seed=seeds[r]) for r in range(R)]
    types = [t0.copy() for _ in range(R)]
    seen_states = set()
    for k, pl in enumerate(plans):
        eng.run(k, 1)
        x, v, t = eng.state()
        fr, hist = eng.frames(k, 1, with_hist=True)
        for r in range(R):
            m = OB.measure(x[r], lay)
            s = machines[r].step(pl.index, types[r], m["n_prom"], int(exp_rnp[pl.index]))
            rec = np.array(OB.frame_record(pl.index, m, s), dtype=np.float64)
            assert np.array_equal(fr[r, 0], rec), (pl.index, r, fr[r, 0], rec)
            assert np.array_equal(t[r], types[r]), (pl.index, r)
            assert list(hist[r, 0]) == OB.shell_histogram(x[r], lay, P.CLUSTER_R)
            seen_states.add((r, s))
    assert {(0, 1), (2, 1), (3, 1)} <= seen_states and (2, 2) not in seen_states and (3, 2) not in seen_states
    if promoter == 3:
        assert (0, 2) in seen_states
    assert eng.counters()["total_active"][0] == machines[0].total_active
    # the quirk itself: last promoter bead (p = 3), gene start bead (p = 2), second gene bead (p = 1)
    assert lay.probe - box11.gene_start() == {3: -1, 2: 0, 1: 1}[promoter]
    assert np.all(eng.status() == 0)
    eng.close()


def test_frame_times_and_units(box11, snapshots):
    eng = E.ShoeboxBatch(box11, n_replicas=1, seeds=[1])
    eng.set_schedule(3, observe=True, n_steps=5)
    eng.run()
    fr = eng.frames()
    assert list(fr[0, :, 0]) == [6.0, 12.0, 18.0]                    # (i+1) * delt, delt = 0.03 * tRun
    x, _, _ = eng.state()
    g = box11.gene_start()
    assert np.array_equal(fr[0, 2, 1:4], x[0, g - 1] * 60)           # probe bead (promoter length 3), nm
    eng.close()
