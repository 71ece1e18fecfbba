"""Level 3: ensemble statistics of the device path against the fp64 oracle (statistical parity).

Reference ensembles: tests/golden/oracle_ensemble_box11.npz = 48 oracle replicas x 400 chunks and
oracle_ensemble_box11_2000.npz = 48 x 2000 chunks, the production run length
(box 11, promoter 3, -x 20, -a 100; made by tests/golden/make_oracle_ensemble.py).  The device runs
192 replicas of the same protocol with different seeds.  Window means must agree within 4.5
standard errors of the difference (both ensembles contribute); quantities fixed by the protocol
(RNP count, frame times) must agree exactly.  Replicas that fail on the way (about 1 in 1000 production runs does) are
dropped, as the reference's launcher drops and repeats them."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from enhancershoebox_b200 import datafile, engine as E

pytestmark = pytest.mark.gpu
WINDOWS = ((20, 100), (100, 200), (200, 300), (300, 400))
WINDOWS_FULL = ((20, 150), (150, 400), (400, 800), (800, 1200), (1200, 1600), (1600, 2000))


def _sem_diff(a, b):
    return np.sqrt(a.var(ddof=1) / len(a) + b.var(ddof=1) / len(b))


@pytest.mark.parametrize("fixture,windows", [("oracle_ensemble_box11.npz", WINDOWS), ("oracle_ensemble_box11_2000.npz", WINDOWS_FULL)])
def test_ensemble_matches_oracle(fixture, windows):
    """400 chunks (12 % RNP at the end) and the production run length, 2000 chunks x 200 steps (90 % RNP, where the
    cut-2.9 RNP-RNP attraction dominates): fp32 state on the device over 4e5 steps against the fp64 oracle."""
    z = np.load(os.path.join(GOLDEN, fixture))
    nch, thr, act = [int(v) for v in z["meta"]]
    R = 192
    datas = [datafile.make_shoebox(11, 3, False, seed=10_000 + r) for r in range(R)]
    eng = E.ShoeboxBatch(datas, seeds=np.arange(R) + 777, thresholds=thr, activations=act)
    eng.set_schedule(nch, observe=True)
    eng.run()
    # A production run fails now and then ("Pair distance < table inner cutoff" after a type change inside the cluster:
    # 5-7 of 6600 runs in the committed 330-cell sweeps, profiles/r2_reference_ensemble_all*.md) and the reference's
    # launcher deletes and re-runs such a run (run_parallel_shoebox.sh:1-14): failed replicas are left out here, and
    # there must be few of them.
    ok = eng.status() == 0
    assert (~ok).sum() <= max(2, R // 64) and np.all(z["status"] == 0), eng.status()[~ok]     # (192 runs at ~1.3e-3 each: 3 or more once in 500 test runs)
    fr = eng.frames()[ok]
    x, v, t = (q[ok] for q in eng.state())
    R = int(ok.sum())
    # exact, protocol-determined quantities
    assert np.array_equal(fr[:, :, 0], np.tile((np.arange(nch) + 1) * 6.0, (R, 1)))
    assert np.all((t == 5).sum(axis=1) == z["nrnp"][0, -1])
    assert np.all(fr[:, :150, 8] == 0) and np.all(fr[:, 150, 8] >= 1)
    # temperature: <KE>/N = 3/2 kT (N - 4 anchors)/N
    ke = 0.5 * (v ** 2).sum(axis=(1, 2)) / x.shape[1]
    assert abs(ke.mean() - z["ke"][:, 20:].mean()) < 0.02
    report = []
    for lo, hi in windows:
        for name, col in (("d_rp", 4), ("d_rg", 5), ("n_prom", 6), ("n_gene", 7)):
            g = fr[:, lo:hi, col].mean(axis=1)
            o = z[name][:, lo:hi].astype(np.float64).mean(axis=1)
            zscore = (g.mean() - o.mean()) / _sem_diff(g, o)
            report.append((lo, hi, name, g.mean(), o.mean(), zscore))
            assert abs(zscore) < 4.5, report[-1]
        ga = (fr[:, lo:hi, 8] == 2).mean(axis=1)
        oa = (z["state"][:, lo:hi] == 2).mean(axis=1)
        if lo >= 100:
            assert abs(ga.mean() - oa.mean()) < 4.5 * _sem_diff(ga, oa) + 1e-3, (lo, hi, ga.mean(), oa.mean())
    print("\n".join(f"[{lo},{hi}) {n}: device {g:.2f} oracle {o:.2f} z {zs:+.2f}" for lo, hi, n, g, o, zs in report))
    eng.close()
