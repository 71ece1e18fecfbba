"""The committed device ensembles (profiles/r2_reference_ensemble_*.json and, with the final round-2 kernel,
r3_reference_ensemble_*.json, written on a B200 by
scripts/reference_ensemble.py) against the reference's own LAMMPS sweep -- the same comparison as
tests/test_gpu_reference_ensemble.py, re-checkable without a GPU -- and the known answers of the golden table."""
import glob
import json
import os

import numpy as np
import pytest

import ensemble_compare as EC

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_golden_table_known_answers():
    cells, rows = EC.load_golden()
    assert len(cells) == 330 and all(c["n"] == 100 for c in cells.values())
    c = cells[(3, 80, 30)]                                  # 17.28 +- 19.80, 6.67 +- 9.97, 7.19 +- 10.86, 123.2 +- 19.5 (44)
    assert abs(c["s5p"][0] - 17.28) < 0.01 and abs(c["s2p"][0] - 6.67) < 0.01 and abs(c["contact"][0] - 7.19) < 0.01
    assert c["n_act"] == 44 and abs(c["dist"][0] - 123.2) < 0.1
    for r in rows.values():                                 # the S2PInt quantum of the reference's own runs
        q = r[:, 1] / EC.S2P_QUANTUM
        assert np.allclose(q, np.round(q), atol=1e-6)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(ROOT, "profiles", "r[23]_reference_ensemble_*.json"))))
def test_committed_device_ensembles_match_the_reference(path):
    """Stated tolerances.  Cells with the reference's own 100 repeats: |z| < 4 per comparison.  The 330-cell sweeps at 20
    repeats per cell (heavy-tailed per-run distributions, small samples): |z| < 6.5, at most 3 % of the comparisons beyond
    3, mean z over all comparisons within 0.25 of zero.  Per-promoter means of S5PInt, S2PInt and Contact (BASELINE.md's
    last row): the reference's launcher starts every repeat of a promoter length from ONE bundled data file, and the table
    is conditioned on that configuration -- the sweep that starts from the same files (`ic: bundled`) must agree within 2.5
    combined standard errors and 8 %; sweeps from fresh make_input_file configurations sit up to 15 % off (promoter 2 and
    3: +10 %, the finding that led to the bundled-IC run) and are only required to stay within 20 %."""
    comp, res = EC.compare_result_file(path)
    cells, _ = EC.load_golden()
    assert res["chunks"] == 2000 and res["box"] == 11
    n_runs = sum(len(v) for v in res["cells"].values())
    assert res["failed"] <= 0.03 * (n_runs + res["failed"])
    zs = np.array([q[-1] for c in comp.values() for q in c if np.isfinite(q[-1])])
    if res["repeats"] >= 100:
        assert np.all(np.abs(zs) < 4.0), EC.table(comp)
    else:
        assert np.all(np.abs(zs) < 6.5) and np.mean(np.abs(zs) > 3.0) < 0.03, (np.abs(zs).max(), np.mean(np.abs(zs) > 3.0))
    assert abs(zs.mean()) < 0.25 + 4.0 / np.sqrt(len(zs)) and 0.5 < zs.std() < 1.5, (zs.mean(), zs.std())
    if len(comp) == 330:
        pm = EC.promoter_means(res, cells)
        for p, rows in pm.items():
            for name, dm, dse, rm, rse, z in rows:
                if res.get("ic") == "bundled":
                    assert abs(z) < 2.5 and abs(dm / rm - 1.0) < 0.08, (p, name, dm, rm, z)
                else:
                    assert abs(dm / rm - 1.0) < 0.20, (p, name, dm, rm, z)
