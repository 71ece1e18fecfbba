"""configs[2] in miniature: the sweep runner (conditions -> replicas -> device-side 5-run summary rows) against the
same runs read back to the host and reduced by the numpy restatement of dist_genestages_grouped.py."""
import numpy as np
import pytest

from enhancershoebox_b200 import datafile, engine as E, sweep
from oracle import aggregate as AGG

pytestmark = pytest.mark.gpu


def test_sweep_rows_equal_host_aggregation_of_the_same_runs():
    box, nch = 9, 170
    groups, rows = sweep.run_sweep(box, promoters=(3, 2), activations=(100,), thresholds=(10, 20), repeats=5, n_chunks=nch, batch_groups=2)
    assert groups == [(3, 10, 100, 1), (3, 20, 100, 1), (2, 10, 100, 1), (2, 20, 100, 1)] and rows.shape == (4, 4)
    for g, (p, t, a, _) in enumerate(groups):
        seeds = np.array([g * 5 + k + 1 for k in range(5)], dtype=np.uint64)
        datas = [datafile.make_shoebox(box, p, False, seed=int(s)) for s in seeds]
        eng = E.ShoeboxBatch(datas, seeds=seeds, thresholds=t, activations=a)
        eng.set_schedule(nch, observe=True)
        eng.run()
        assert np.all(eng.status() == 0)
        ref = AGG.grouped(eng.frames())
        eng.close()
        assert np.array_equal(rows[g], ref[0], equal_nan=True), (g, rows[g], ref[0])
    assert np.all(rows[:, 0] >= 0) and np.all((rows[:, 2] >= 0) & (rows[:, 2] <= 100))
    csv = sweep.summary_csv(groups, rows)
    assert len(csv.splitlines()) == 5


def test_run_sweep_two_ranks_nccl(tmp_path):
    """run_sweep.py as its own docstring launches it -- torchrun, one rank per GPU, process group "nccl" -- on two GPUs:
    the end-of-run all_gather of the per-group rows must move DEVICE tensors (a CPU tensor has no NCCL backend) and the
    table must equal the one-rank table row for row.  Skipped where fewer than two GPUs are visible."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    common = ["-b", "9", "--promoters", "3", "2", "--activations", "100", "--thresholds", "10", "20", "--repeats", "5", "-t", "170", "--batch-groups", "2"]
    one = os.path.join(tmp_path, "one.txt")
    two = os.path.join(tmp_path, "two.txt")
    subprocess.check_call([sys.executable, os.path.join(root, "run_sweep.py"), *common, "-o", one], cwd=root)
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                           "--master-port", "29611", os.path.join(root, "run_sweep.py"), *common, "-o", two], cwd=root)
    assert open(one).read() == open(two).read() and len(open(two).read().splitlines()) == 5
