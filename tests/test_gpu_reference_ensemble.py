"""Level 3 parity against the reference-held ensemble tables (VisitorGeneMaps/
summary_contact_all_Thresholds10-100.txt is the statistical oracle; BASELINE.md quotes five cells).

The device runs the PRODUCTION protocol -- box 11, the reference's bundled initial condition of each promoter length
(tests/golden/bundled_ic_bs11.npz: the reference starts all repeats of a promoter length from that one file), soft phase, table
switch, 2000 Python timesteps x 200 MD steps, gene state machine, RBP->RNP ramp to 90 % -- 100 repeats per cell for
the five quoted cells and compares S5PInt, S2PInt, Contact, DistActivation (mean over repeats; DistActivation over the
activating repeats) and the fraction of activating repeats with the 100 LAMMPS repeats of the same cell.
Stated tolerance: |z| < 4 per comparison (both samples' standard errors); S5PInt, S2PInt, Contact and the activating
fraction of one cell move together (all four follow whether the Ser5P cluster sits at the gene), so the joint statistic
takes two weakly correlated quantities per cell: sum over the five cells of z(S5PInt)^2 + z(DistActivation)^2 < 35.6
(chi^2_10: p = 1e-4); S2PInt a whole number of frames out of 1850.  The 330-cell sweep (profiles/
r2_reference_ensemble_all.md, checked by tests/test_reference_ensemble.py) is the test of systematic deviations."""
import dataclasses
import os

import numpy as np
import pytest

import ensemble_compare as EC
from conftest import GOLDEN
from enhancershoebox_b200 import datafile, engine as E

pytestmark = pytest.mark.gpu
QUOTED = [(3, 80, 30), (1, 80, 30), (2, 50, 10), (3, 10, 100), (3, 100, 1)]
REPEATS = 100


def test_production_ensembles_match_the_reference_tables():
    cells, _ = EC.load_golden()
    ic = np.load(os.path.join(GOLDEN, "bundled_ic_bs11.npz"))
    zs, joint, report = [], [], {}
    for p in (1, 2, 3):
        runs = [(c, k) for c in QUOTED if c[0] == p for k in range(REPEATS)]
        seeds = np.array([31_000_000 + 1000 * (c[1] * 1000 + c[2]) + k for c, k in runs], dtype=np.uint64)
        # every repeat starts from the reference's ONE bundled data file of this promoter length, as the reference's launcher does
        d0 = dataclasses.replace(datafile.make_shoebox(11, p, False, seed=0), x=ic[f"x{p}"].astype(np.float64), types=ic[f"t{p}"].astype(np.int32))
        eng = E.ShoeboxBatch(d0, n_replicas=len(runs), seeds=seeds, thresholds=[c[1] for c, _ in runs], activations=[c[2] for c, _ in runs])
        eng.set_schedule(2000, observe=True)
        eng.run()
        ok = eng.status() == 0
        assert ok.mean() > 0.97                       # LAMMPS loses a run now and then too (the launcher retries)
        rows = E.aggregate_all(eng.frames())
        eng.close()
        for c in QUOTED:
            if c[0] != p:
                continue
            sel = np.array([cc == c for cc, _ in runs]) & ok
            comp = EC.compare_cell(rows[sel], cells[c])
            report[",".join(map(str, c))] = comp
            assert EC.s2p_is_quantised(rows[sel])
            zs += [q[-1] for q in comp]
            joint += [q[-1] for q in comp if q[0] in ("S5PInt", "DistActivation")]
    print(EC.table(report))
    zs, joint = np.array(zs), np.array(joint)
    assert np.all(np.abs(zs) < 4.0), EC.table(report)
    assert float((joint ** 2).sum()) < 35.6, EC.table(report)
