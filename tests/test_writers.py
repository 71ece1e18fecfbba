"""(f-1) on-disk writers and the drop-in command line (host logic, no GPU)."""
import os

import numpy as np

from enhancershoebox_b200 import drivers, protocol as P, writers


def test_event_log_windows():
    st = np.zeros(200)
    st[100:150] = 2
    ev = writers.event_log(st)
    assert ev[100] != 2 and ev[101] == 2 and ev[150] == 2 and ev[151] != 2       # shifted by one frame (:1596-1598)
    assert all(e == 1 for e in ev[81:101]) and ev[80] == 0                        # approaching: w1 = 20 frames before
    assert all(e == 3 for e in ev[151:171]) and ev[171] == 0                      # receding: w2 = 20 frames after


def test_write_run_files(tmp_path):
    n = 400
    rng = np.random.default_rng(1)
    frames = np.zeros((n, 9))
    frames[:, 0] = (np.arange(n) + 1) * 6.0
    frames[:, 1:4] = rng.normal(size=(n, 3)) * 60
    raw = rng.uniform(1, 8, size=(n, 2))
    frames[:, 4], frames[:, 5] = raw[:, 0] * 60, raw[:, 1] * 60
    frames[:, 6:8] = rng.integers(0, 40, size=(n, 2))
    frames[150:, 8] = 1
    frames[200:250, 8] = 2
    hist = rng.integers(0, 9, size=(n, 49))
    out = str(tmp_path / "box11" / "Control_Promoter3_Threshold80_Act30")
    writers.write_run(out, 7, frames, hist, raw, 50)
    run = os.path.join(out, "run7")
    lines = open(os.path.join(run, "geneTrack.txt")).read().split("\n")
    assert len(lines) == n + 1 and lines[0].startswith("6.0,") and lines[0].count(",") == 8
    back = np.array([[float(v) for v in l.split(",")] for l in lines[:-1]])
    assert np.array_equal(back, frames)                                           # str(float) round-trips exactly
    assert lines[0].split(",")[6:] == [str(int(frames[0, 6])), str(int(frames[0, 7])), "0"]
    cl = open(os.path.join(run, "ser5pAroundCluster.txt")).read().split("\n")
    assert cl[0] == "t," + ",".join(str(v) for v in P.CLUSTER_R[1:]) and len(cl[1].split(",")) == 50 and len(cl) == n + 2
    assert open(os.path.join(out, "parallel_counter", "run7.txt")).read() == "Run 7 done!\n"
    assert open(os.path.join(out, "active_duration.txt")).read() == "7,50,1850\n"
    gs = open(os.path.join(out, "gene_stats.txt")).read().split("\n")[:-1]
    assert len(gs) == n - 150
    first = gs[0].split(",")
    assert first[0] == ("1" if raw[150, 1] < 3 else "0") and float(first[1]) == raw[150, 1] * 60 and first[3] == "0"
    assert sum(int(l.split(",")[3]) for l in gs) == 50                             # frames 201..250 carry event 2
    total = sum(len(open(os.path.join(out, f"dist_{e}.txt")).read().split("\n")) - 1 for e in writers.EVENT_TYPES)
    assert total == n
    total_dd = sum(len(open(os.path.join(out, f"ddist_{e}.txt")).read().split("\n")) - 1 for e in writers.EVENT_TYPES)
    assert total_dd == n - 1
    assert os.path.exists(os.path.join(run, "figures", "ser5p_cluster.pdf"))       # the launcher's completion sentinel
    for sub in ("figures", "image_files", "microscopy_files"):
        assert os.path.isdir(os.path.join(run, sub))


def test_cli_flags():
    a = drivers.parse_args("-b 12 -r 3 -t 100 -o out/x/ -m 0 -c LatB -p 2 -a 30 -x 80 -z 300 -n 0.001".split(), with_actin=True)
    assert (a["box"], a["repeats"], a["total"], a["out"], a["condition"], a["promoter"], a["activation"], a["threshold"], a["actin"]) == \
        (12, [3], 100, "out/x/", "LatB", 2, 30.0, 80, 300)
    b = drivers.parse_args("--box 11 --repeat 1:8 --promoter 3 --activation 5 --threshold 20".split(), with_actin=False)
    assert b["repeats"] == list(range(1, 9)) and b["activation"] == 5.0
