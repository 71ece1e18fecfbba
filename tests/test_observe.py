"""a14-a19 on the CPU side: measurement restatement, gene state machine, aggregation oracle."""
import io
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import aggregate as AGG
from oracle import observe as OB
from enhancershoebox_b200 import protocol as P


def test_probe_index_quirk(box11):
    # atomPositions[x - promoter_length + 1] with x the 1-based gene start
    from enhancershoebox_b200 import datafile
    for p, off in ((3, -1), (2, 0), (1, 1)):
        d = datafile.make_shoebox(11, p, False, seed=1)
        lay = OB.Layout.from_data(d)
        assert lay.probe == d.gene_start() + off
        assert lay.promoter_rows == [d.gene_start() - 1 - k for k in range(p)]


def test_measure_matches_direct_expressions(box11, snapshots):
    lay = OB.Layout.from_data(box11)
    x = snapshots["x500"].astype(np.float64)
    m = OB.measure(x, lay)
    rr = x[lay.enhancer].sum(axis=0) / 50
    g = lay.gene_start
    assert abs(m["d_rg"] - np.linalg.norm(rr - x[g])) < 1e-12
    assert abs(m["d_rp"] - np.linalg.norm(rr - x[[g - 1, g - 2, g - 3]].mean(axis=0))) < 1e-12
    d = np.linalg.norm(x[lay.ser5p] - x[g - 1], axis=1)
    assert m["n_prom"] == int((d < 3).sum())
    assert m["n_gene"] == int((np.linalg.norm(x[lay.ser5p] - x[g], axis=1) < 3).sum())
    h = OB.shell_histogram(x, lay, P.CLUSTER_R)
    dmin = np.min(np.linalg.norm(x[lay.ser5p][:, None, :] - x[lay.enhancer][None], axis=2), axis=1)
    assert h == list(np.histogram(dmin, bins=P.CLUSTER_R)[0]) or sum(h) == int(((dmin >= 0.65) & (dmin < 2.5)).sum())
    assert len(h) == 49


def test_numpy_axis0_mean_is_a_sequential_sum(snapshots):
    """The device reproduces np.mean(x[idx], axis=0) as a left-to-right sum followed by one division."""
    x = snapshots["x1999"].astype(np.float64)[37:87]
    s = np.zeros(3)
    for row in x:
        s = s + row
    assert np.array_equal(np.mean(x, axis=0), s / 50)


def test_gene_machine_timeline(box11):
    """Strict inequalities, activation delay 2, active window 50."""
    lay = OB.Layout.from_data(box11)
    gm = OB.GeneMachine(lay, threshold=10, p_activation=1.0, seed=5)
    types = box11.types.copy()
    g, p = lay.gene_start, 3
    states = []
    for i in range(260):
        n_prom = 11 if i >= 160 else 10          # count == threshold must NOT activate (needs count > threshold)
        states.append(gm.step(i, types, n_prom, -1))
        if i == 149:
            assert np.all(types[g:g + 5] == 2)
        if i == 150:
            assert np.all(types[g:g + 5] == 6)  # induced as soon as i >= 150
        if i == 161:
            assert np.all(types[g:g + 5] == 6) and np.all(types[g - p:g] == 10)   # marked, beads not yet flipped
        if i == 162:
            assert np.all(types[g:g + 5] == 7) and np.all(types[g - p:g] == 11)   # flipped after activation_delay = 2
    s = np.array(states)
    assert np.all(s[:150] == 0) and np.all(s[150:160] == 1)
    assert np.all(s[160:210] == 2) and s[210] == 0 and s[211] == 2    # 50 frames active, one frame 0, re-induced+re-marked
    assert gm.total_active == (s == 2).sum() - 2 or gm.total_active > 0


def test_gene_machine_probability_gate(box11):
    lay = OB.Layout.from_data(box11)
    never = OB.GeneMachine(lay, threshold=0, p_activation=0.0, seed=9)
    types = box11.types.copy()
    assert all(never.step(i, types, 50, -1) < 2 for i in range(150, 400))
    hits = []
    for seed in range(40):
        gm = OB.GeneMachine(lay, threshold=0, p_activation=0.05, seed=seed)
        t = box11.types.copy()
        first = next((i for i in range(150, 600) if gm.step(i, t, 5, -1) == 2), None)
        hits.append(first - 150 if first is not None else 450)
    assert 8 < np.mean(hits) < 40                                   # geometric waiting time, mean 1/p = 20


def test_rnp_ramp(box11):
    lay = OB.Layout.from_data(box11)
    exp = OB.expected_rnp_table(2000, lay.n_rbprnp, 2000)
    assert np.all(exp[:149] == -1) and exp[149] == 0 and exp[1999] == 495     # ~495 of 550 at the end
    assert np.all(np.diff(exp[149:]) >= 0) and np.all(np.diff(exp[149:]) <= 1)
    gm = OB.GeneMachine(lay, threshold=10 ** 6, p_activation=0.0, seed=77)
    types = box11.types.copy()
    for i in range(149, 400):
        gm.step(i, types, 0, int(exp[i]))
        assert (types == 5).sum() == exp[i] and (types == 4).sum() + (types == 5).sum() == 550
    gm2 = OB.GeneMachine(lay, threshold=10 ** 6, p_activation=0.0, seed=78)
    types2 = box11.types.copy()
    for i in range(149, 400):
        gm2.step(i, types2, 0, int(exp[i]))
    assert not np.array_equal(types, types2)                          # which RBPs convert depends on the seed


def test_frame_line_format():
    rec = OB.frame_record(0, dict(probe=np.array([0.5, -1.25, 2.0]), d_rp=3.0, d_rg=4.5, n_prom=7, n_gene=9), 1)
    assert OB.frame_line(rec) == "6.0,30.0,-75.0,120.0,180.0,270.0,7,9,1\n"


def _golden_frames(z, key):
    d_rp, d_rg, s5p, st = z["drp_" + key], z["drg_" + key], z["s5p_" + key], z["state_" + key]
    fr = np.zeros(d_rp.shape + (9,))
    fr[:, :, 4], fr[:, :, 5], fr[:, :, 6], fr[:, :, 8] = d_rp, d_rg, s5p, st
    return fr


def test_aggregate_oracle_equals_reference_script():
    """oracle/aggregate.py against the CSV written by the unmodified reference script (bit-exact: the
    CSV holds repr() of the float64 values)."""
    import pandas as pd
    z = np.load(os.path.join(GOLDEN, "aggregate_golden.npz"))
    df = pd.read_csv(io.StringIO(str(z["csv"])), index_col=0, float_precision="round_trip")
    assert len(df) == 5
    for (p, t, a) in ((1, 10, 1), (2, 50, 10), (3, 80, 30)):
        rows = AGG.grouped(_golden_frames(z, f"{p}_{t}_{a}"))
        ref = df[(df.Promoter == p) & (df.Threshold == t) & (df.Activation == a)][["S5PInt", "S2PInt", "Contact", "DistActivation"]].to_numpy(float)
        assert rows.shape == ref.shape
        assert np.array_equal(rows, ref, equal_nan=True), (p, t, a)
    assert np.isnan(df[df.Promoter == 1].DistActivation).all()        # groups without any activation -> NaN


def test_percentile_oracle_equals_reference_script():
    """oracle/aggregate.percentile_10x against the CSV of the unmodified
    dist_genestages_grouped_5percentile_10xaveraging.py (groups of 10 + the remainder row)."""
    import pandas as pd
    z = np.load(os.path.join(GOLDEN, "aggregate_golden.npz"))
    df = pd.read_csv(io.StringIO(str(z["csv_pct"])), float_precision="round_trip")
    assert len(df) == 3
    cols = ["S5PInt", "S2PInt", "Contact", "DistActivation", "5percentRG", "5percentRP"]
    for (p, t, a) in ((1, 10, 1), (2, 50, 10), (3, 80, 30)):
        labels, rows = AGG.percentile_10x(_golden_frames(z, f"{p}_{t}_{a}"))
        sel = df[(df.Promoter == p) & (df.Threshold == t) & (df.Activation == a)]
        assert labels == list(sel.Runs)
        assert np.array_equal(rows, sel[cols].to_numpy(float), equal_nan=True), (p, t, a)


def test_reference_summary_tables_quantum():
    """Known answer on the shipped tables: S2PInt per run is a multiple of 100/1850."""
    path = "/root/reference/VisitorGeneMaps/summary_contact_all_Thresholds10-100.txt"
    if not os.path.exists(path):
        pytest.skip("reference checkout not present on this machine")
    import pandas as pd
    df = pd.read_csv(path, index_col=0)
    assert len(df) == 33000
    q = df.S2PInt.to_numpy() / (100 / 1850)
    assert np.mean(np.abs(q - np.round(q)) < 1e-6) > 0.94          # 31119 of 33000 rows are exact multiples (filter_ser2p_points.py:38-41)
