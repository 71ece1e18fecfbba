"""a19: the aggregation kernel is bit-exact with the reference script (golden made by running the
unmodified VisitorGeneMaps/dist_genestages_grouped.py) and with the numpy restatement."""
import io
import os

import numpy as np
import pytest

from conftest import GOLDEN
from enhancershoebox_b200 import engine as E
from oracle import aggregate as AGG

pytestmark = pytest.mark.gpu


def _golden_frames(z, key):
    d_rp, d_rg, s5p, st = z["drp_" + key], z["drg_" + key], z["s5p_" + key], z["state_" + key]
    fr = np.zeros(d_rp.shape + (9,))
    fr[:, :, 4], fr[:, :, 5], fr[:, :, 6], fr[:, :, 8] = d_rp, d_rg, s5p, st
    return fr


def test_against_reference_script_output():
    import pandas as pd
    z = np.load(os.path.join(GOLDEN, "aggregate_golden.npz"))
    df = pd.read_csv(io.StringIO(str(z["csv"])), index_col=0, float_precision="round_trip")
    for (p, t, a) in ((1, 10, 1), (2, 50, 10), (3, 80, 30)):
        rows = E.aggregate_grouped(_golden_frames(z, f"{p}_{t}_{a}"))
        ref = df[(df.Promoter == p) & (df.Threshold == t) & (df.Activation == a)][["S5PInt", "S2PInt", "Contact", "DistActivation"]].to_numpy(float)
        assert np.array_equal(rows, ref, equal_nan=True), (p, t, a)


@pytest.mark.parametrize("n_runs,n_frames", [(5, 151), (10, 2000), (15, 777), (5, 10000), (7, 300)])
def test_random_frames_against_numpy_restatement(n_runs, n_frames):
    rng = np.random.default_rng(n_runs * 100003 + n_frames)
    fr = np.zeros((n_runs, n_frames, 9))
    fr[:, :, 4] = rng.uniform(50, 900, (n_runs, n_frames))
    fr[:, :, 6] = rng.integers(0, 120, (n_runs, n_frames))
    st = rng.integers(0, 3, (n_runs, n_frames))
    st[:, : n_frames // 3] = 0
    fr[:, :, 8] = st
    out = E.aggregate_grouped(fr)
    ref = AGG.grouped(fr)
    assert out.shape == (n_runs // 5, 4) and np.array_equal(out, ref, equal_nan=True)


def test_config5_256_runs_of_10000_frames():
    """configs[4]: the analysis alone on 256 runs x 1e4 frames (synthetic frames as SURVEY 8d: d_rp, d_rg ~ U(50, 900) nm,
    s5p ~ randint(0, 120), a 3-state script), bit-exact against the Python statements of the reference script."""
    import time
    rng = np.random.default_rng(0)
    n_runs, n_frames = 256, 10000
    fr = np.zeros((n_runs, n_frames, 9))
    fr[:, :, 4] = rng.uniform(50, 900, (n_runs, n_frames))
    fr[:, :, 5] = rng.uniform(50, 900, (n_runs, n_frames))
    fr[:, :, 6] = rng.integers(0, 120, (n_runs, n_frames))
    state = np.zeros((n_runs, n_frames), np.int64)
    for r in range(n_runs):                                   # inactive -> induced at 150 -> bursts of 50 active frames
        state[r, 150:] = 1
        for start in rng.integers(160, n_frames - 60, rng.integers(0, 12)):
            state[r, start:start + 50] = 2
    fr[:, :, 8] = state
    t0 = time.time()
    out = E.aggregate_grouped(fr)
    t_gpu = time.time() - t0
    ref = AGG.grouped(fr)
    assert out.shape == (51, 4) and np.array_equal(out, ref, equal_nan=True)
    rows = E.aggregate_runs(fr)
    assert np.array_equal(rows, AGG.per_run_percentile(fr), equal_nan=True)
    print(f"256 x 1e4 frames: {t_gpu * 1e3:.1f} ms incl. the 184 MB host->device copy")


def test_ragged_and_empty_cases():
    fr = np.zeros((5, 400, 9))
    fr[:, :, 4] = 300.0                                     # never in contact, never active
    out = E.aggregate_grouped(fr)
    assert out[0, 0] == 0.0 and out[0, 1] == 0.0 and out[0, 2] == 0.0 and np.isnan(out[0, 3])
    assert E.aggregate_grouped(np.zeros((4, 400, 9))).shape == (0, 4)        # fewer runs than one group: no row
    fr[:, :, 8] = 2.0                                      # active from frame 0: frame 0 is never an onset (I, J start at 1)
    fr[:, :, 4] = 100.0
    out = E.aggregate_grouped(fr)
    assert out[0, 1] == 100 * 399 / 250 and out[0, 2] == 100.0 and np.isnan(out[0, 3])


def test_percentile_variant_against_reference_script_output():
    """(f)3: per-run statistics + 5th percentiles on the device, 10x nanmean grouping on the host, against the CSV
    of the unmodified dist_genestages_grouped_5percentile_10xaveraging.py."""
    import pandas as pd
    z = np.load(os.path.join(GOLDEN, "aggregate_golden.npz"))
    df = pd.read_csv(io.StringIO(str(z["csv_pct"])), float_precision="round_trip")
    cols = ["S5PInt", "S2PInt", "Contact", "DistActivation", "5percentRG", "5percentRP"]
    for (p, t, a) in ((1, 10, 1), (2, 50, 10), (3, 80, 30)):
        labels, rows = E.aggregate_percentile_10x(_golden_frames(z, f"{p}_{t}_{a}"))
        sel = df[(df.Promoter == p) & (df.Threshold == t) & (df.Activation == a)]
        assert labels == list(sel.Runs)
        assert np.array_equal(rows, sel[cols].to_numpy(float), equal_nan=True), (p, t, a)


@pytest.mark.parametrize("n_runs,n_frames,q", [(3, 152, 5), (4, 2000, 5), (2, 10000, 5), (3, 777, 50), (2, 1000, 99.5), (2, 400, 0), (2, 400, 100)])
def test_per_run_rows_against_numpy(n_runs, n_frames, q):
    rng = np.random.default_rng(n_runs * 7919 + n_frames)
    fr = np.zeros((n_runs, n_frames, 9))
    fr[:, :, 4] = rng.uniform(50, 900, (n_runs, n_frames))
    fr[:, :, 5] = np.round(rng.uniform(50, 900, (n_runs, n_frames)))         # ties
    fr[:, :, 6] = rng.integers(0, 120, (n_runs, n_frames))
    st = rng.integers(0, 3, (n_runs, n_frames))
    st[:, : n_frames // 3] = 0
    st[0] = 0                                                                  # a run that never activates -> NaN
    fr[:, :, 8] = st
    out = E.aggregate_runs(fr, q=q)
    assert np.array_equal(out, AGG.per_run_percentile(fr, q=q), equal_nan=True)
