"""(e) sharding + end-of-run gather; world_size 2 over gloo on CPU."""
import os
import socket

import numpy as np
import pytest

from enhancershoebox_b200 import sweep


def test_shard_ranges_cover_everything():
    for n, w in ((4096, 8), (2304, 8), (64, 1), (10, 4), (3, 8)):
        spans = [sweep.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        c = sweep.shard_counts(n, w)
        assert sum(c) == n and max(c) - min(c) <= 1
    assert sweep.shard_counts(4096, 8) == [512] * 8
    with pytest.raises(ValueError):
        sweep.shard_range(10, 4, 4)


def test_sweep_grid_matches_launcher():
    g = sweep.sweep_conditions()
    assert len(g) == 3 * 10 * 11 * 100 == 33000        # rows of summary_contact_all_Thresholds10-100.txt
    assert g[0] == (1, 10, 1, 1) and g[-1] == (3, 100, 100, 100)


def _worker(rank, world, port, n_total, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sweep.shard_range(n_total, rank, world)
    local = np.arange(lo, hi, dtype=np.float64)[:, None] * np.array([1.0, 10.0, 100.0, np.nan])[None, :]
    full = sweep.gather_rows(local, n_total)
    q.put((rank, full))
    dist.destroy_process_group()


def test_gather_rows_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    q = ctx.Queue()
    n_total = 7                                           # uneven shards: 4 + 3
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    exp = np.arange(n_total, dtype=np.float64)[:, None] * np.array([1.0, 10.0, 100.0, np.nan])[None, :]
    for r in range(2):
        assert np.array_equal(got[r], exp, equal_nan=True)


def test_sweep_groups_and_summary_format():
    g = sweep.sweep_groups((1, 2, 3), (1, 5, 10, 15, 20, 25, 30, 50, 75, 100), (10, 20, 30, 40, 50, 60, 70, 75, 80, 90, 100), 100)
    assert len(g) == 6600 and g[0] == (1, 10, 1, 1) and g[19] == (1, 10, 1, 96) and g[20] == (1, 20, 1, 1)      # rows of summary_contact_grouped_*.txt
    csv = sweep.summary_csv(g[:2], np.array([[59.5, 0.0, 23.975, np.nan], [60.25, 8.426395939086294, 24.0, 452.8204550361811]]))
    assert csv.splitlines() == [",Promoter,Threshold,Activation,S5PInt,S2PInt,Contact,DistActivation",
                                "0,1,10,1,59.5,0.0,23.975,", "1,1,10,1,60.25,8.426395939086294,24.0,452.8204550361811"]
    path = "/root/reference/VisitorGeneMaps/summary_contact_grouped_Thresholds10-100.txt"
    if os.path.exists(path):                       # same header and row labels as the shipped table
        ref = open(path).read().splitlines()
        assert ref[0] == csv.splitlines()[0] and len(ref) == 1 + 6600
        assert [r.split(",")[:4] for r in ref[1:21]] == [[str(k), "1", "10", "1"] for k in range(20)]


def test_cpp_format_writer_matches_the_shipped_table():
    """summary_contact_grouped_Thresholds10-100_cpp.txt (written by the reference's C++ aggregator): header without index
    column, 6 significant digits, 0 for a group without activation.  Every row of the shipped file must survive a
    round trip through the writer unchanged (where the reference checkout is present)."""
    import os
    import pytest
    from enhancershoebox_b200 import sweep
    groups = [(1, 10, 1, 1), (3, 80, 30, 6)]
    rows = np.array([[20.125512, 0.54054054054, 17.3, 127.06351], [15.38229, 0.0, 12.15, np.nan]])
    text = sweep.summary_cpp_csv(groups, rows)
    assert text.splitlines() == ["Promoter,Threshold,Activation,S5PInt,S2PInt,Contact,DistActivation",
                                 "1,10,1,20.1255,0.540541,17.3,127.064", "3,80,30,15.3823,0,12.15,0"]
    ref = "/root/reference/VisitorGeneMaps/summary_contact_grouped_Thresholds10-100_cpp.txt"
    if not os.path.exists(ref):
        pytest.skip("reference checkout not present")
    lines = open(ref).read().splitlines()
    assert lines[0] == text.splitlines()[0]
    body = [l.split(",") for l in lines[1:]]
    g = [(int(f[0]), int(f[1]), int(f[2]), 0) for f in body]
    r = np.array([[float(v) for v in f[3:]] for f in body])
    assert sweep.summary_cpp_csv(g, r).splitlines()[1:] == lines[1:]
