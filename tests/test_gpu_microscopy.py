"""(f)4: synthetic-microscopy voxel counts fused on the device, bit-exact against np.histogramdd
(run_single_shoebox.py:152-213, 1498-1502), including the bin-edge rules."""
import os

import numpy as np
import pytest

from enhancershoebox_b200 import datafile, drivers, engine as E, protocol as P, writers
from oracle import observe as OB

pytestmark = pytest.mark.gpu


def test_voxel_counts_equal_histogramdd(box11, snapshots):
    x, v, t = snapshots["x1999"].astype(np.float64), snapshots["v1999"].astype(np.float64), snapshots["t1999"].astype(np.int32)
    R = 3
    xs = np.stack([x] * R)
    # bin-edge cases on free beads: on the first edge (first bin), on the last edge (last bin), just outside, far outside
    # (the z edges +-7 are the walls themselves: a bead there ends the run, so z only gets inner edges)
    xs[1, 900] = [-11.0, 0.0, 0.0]; xs[1, 901] = [11.0, 9.0, 0.0]; xs[1, 902] = [np.nextafter(np.float32(11.0), np.float32(12.0)), 0, 0]
    xs[1, 903] = [0.0, -9.0, 0.0]; xs[1, 904] = [-11.5, 0.0, 0.0]; xs[1, 906] = [0.0, 9.25, 0.0]
    xs[2, 905] = [float(P.MICROSCOPY_EDGES[0][7]), float(P.MICROSCOPY_EDGES[1][3]), float(P.MICROSCOPY_EDGES[2][5])]     # on inner edges
    eng = E.ShoeboxBatch(box11, n_replicas=R, seeds=[1, 2, 3])
    eng.upload_state(xs, np.stack([v] * R), np.stack([t] * R))
    pl = P.chunk_schedule("genes_stages", n_chunks=2000)[1999]                 # (i+1) % 10 == 0
    eng.set_schedule(plans=[P.ChunkPlan(pl.index, pl.pair, None, False, False, False, 0)], observe=True, microscopy=True)
    assert eng.mic_chunks == [1999]
    eng.run()
    xg, _, tg = eng.state()                                                    # zero steps: the uploaded (fp32) positions, updated types
    H = eng.microscopy()
    assert H.shape == (R, 1, 5, 21, 17, 13)
    assert np.all(eng.status() == 0)
    for r in range(R):
        ref = OB.microscopy_histograms(xg[r], tg[r])
        assert np.array_equal(H[r, 0], ref.astype(np.uint16)), r
    assert H[0, 0, 0].sum() > 300 and H[0, 0, 2].sum() > 300                    # Ser5P and chromatin channels are populated
    stack = writers.microscopy_stack(H[0, 0])
    assert stack.shape == (5, 13, 17, 21) and stack.dtype == np.float32
    eng.close()


def test_microscopy_driver_writes_stacks(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    d = datafile.make_shoebox(9, 3, False, seed=3)
    a = drivers.parse_args("-b 9 -r 1:2 -t 2 -o out/ -m 0 -c Control -p 3 -a 30 -x 80".split(), with_actin=False)
    assert drivers.run_condition("genes_stages", a, n_chunks=25, data=d, quiet=True, make_microscopy=True) == 0
    files = sorted(os.listdir(os.path.join("out", "run1", "microscopy_files")))
    assert [f for f in files if f.endswith(".npy")] == ["synthetic_microscopy_0002000.npy", "synthetic_microscopy_0004000.npy"]
    s = np.load(os.path.join("out", "run1", "microscopy_files", "synthetic_microscopy_0004000.npy"))
    assert s.shape == (5, 13, 17, 21) and s[2].sum() > 100                      # chromatin channel: blur conserves the counts inside the volume
