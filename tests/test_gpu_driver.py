"""(b)/(f-1): the drop-in command line end to end on the GPU -- data file in, reference-shaped folders out,
then the aggregation over the written geneTrack.txt files (device kernel == numpy restatement)."""
import os

import numpy as np
import pytest

from enhancershoebox_b200 import datafile, drivers, engine as E
from oracle import aggregate as AGG

pytestmark = pytest.mark.gpu


def test_driver_writes_reference_layout_and_aggregates(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    os.makedirs("input_files")
    d = datafile.make_shoebox(11, 3, False, seed=3)
    path = datafile.input_file_name(11, 3, False)
    assert path == "input_files/IC_ser5P370_RBP550_0RNP_promoter3_bs11.data"      # the name the reference drivers open
    datafile.write_data(path, d)
    out = "box11/Control_Promoter3_Threshold5_Act1000/"
    a = drivers.parse_args(f"-b 11 -r 1:10 -t 10 -o {out} -m 0 -c Control -p 3 -a 1000 -x 5".split(), with_actin=False)
    n_chunks = 230
    assert drivers.run_condition("genes_stages", a, n_chunks=n_chunks, quiet=True) == 0
    frames = []
    for r in range(1, 11):
        run = os.path.join(out, f"run{r}")
        assert os.path.exists(os.path.join(run, "figures", "ser5p_cluster.pdf"))
        assert os.path.exists(os.path.join(out, "parallel_counter", f"run{r}.txt"))
        rows = [[float(v) for v in l.split(",")] for l in open(os.path.join(run, "geneTrack.txt")).read().split("\n")[:-1]]
        assert len(rows) == n_chunks and len(rows[0]) == 9 and rows[0][0] == 6.0
        frames.append(rows)
        assert len(open(os.path.join(run, "ser5pAroundCluster.txt")).read().split("\n")) == n_chunks + 2
    frames = np.array(frames)
    assert np.all(frames[:, :150, 8] == 0) and np.all(frames[:, 150, 8] >= 1)      # induced from i = 150 on
    assert (frames[:, 150:, 8] == 0).sum() <= 10 * 2                              # only the single frame on which a gene is inactivated
    assert (frames[:, :, 8] == 2).sum() > 0                                       # -x 5 -a 1000: genes do activate
    assert len(open(os.path.join(out, "active_duration.txt")).read().split("\n")) == 11
    assert len(open(os.path.join(out, "gene_stats.txt")).read().split("\n")) == 10 * (n_chunks - 150) + 1
    for name in ("dist", "ddist"):
        for ev in ("induced", "approaching", "active", "receding"):
            assert os.path.exists(os.path.join(out, f"{name}_{ev}.txt"))
    rows_gpu = E.aggregate_grouped(frames)
    rows_ref = AGG.grouped(frames)
    assert rows_gpu.shape == (2, 4) and np.array_equal(rows_gpu, rows_ref, equal_nan=True)
    # two processes with the same ESB_SEED write the same bytes
    out2 = "again/"
    a2 = dict(a, out=out2, repeats=[3])
    assert drivers.run_condition("genes_stages", a2, n_chunks=n_chunks, quiet=True) == 0
    assert open(os.path.join(out2, "run3", "geneTrack.txt")).read() == open(os.path.join(out, "run3", "geneTrack.txt")).read()


def test_failed_repeat_is_rerun_with_a_new_seed(tmp_path, monkeypatch):
    """The reference launcher deletes the folder of a repeat whose LAMMPS run died and runs it again
    (run_parallel_shoebox.sh:1-14, 69-75); the batch driver redoes the replicas whose status is non-zero."""
    monkeypatch.chdir(tmp_path)
    d = datafile.make_shoebox(9, 3, False, seed=3)
    calls = []

    class Flaky(E.ShoeboxBatch):
        def status(self):
            st = super().status().copy()
            calls.append(len(st))
            if len(calls) == 1:
                st[1] = E.ST_WALL                      # pretend repeat 2 lost an atom in the first attempt
            return st

    monkeypatch.setattr(drivers, "ShoeboxBatch", Flaky)
    a = drivers.parse_args("-b 9 -r 1:3 -t 3 -o out/ -m 0 -c Control -p 3 -a 30 -x 80".split(), with_actin=False)
    assert drivers.run_condition("genes_stages", a, n_chunks=12, data=d, quiet=True) == 0
    assert calls == [3, 1]                             # second batch: only the failed repeat
    for r in (1, 2, 3):
        assert os.path.exists(os.path.join("out", f"run{r}", "figures", "ser5p_cluster.pdf"))
        assert len(open(os.path.join("out", f"run{r}", "geneTrack.txt")).read().splitlines()) == 12


def test_run_refuses_a_chunk_whose_pair_setup_was_never_given(box11):
    """A schedule that names table map "T" (treatment) on a Control batch, for which only maps A and B were uploaded:
    esb_run must fail (ESB_ESTATE) instead of stepping without pair forces."""
    from enhancershoebox_b200 import engine as E, protocol as P
    eng = E.ShoeboxBatch(box11, n_replicas=2, seeds=[1, 2])
    plans = [P.ChunkPlan(p.index, "T", p.ramp, p.adapt_soft, p.adapt_bond1, p.adapt_bond2, 5) for p in P.chunk_schedule("genes_stages", n_chunks=22)[20:]]
    eng.set_schedule(plans=plans, observe=False)
    with pytest.raises(E.EsbError, match="-4"):
        eng.run()
    eng.close()


def test_unchanged_launcher_spreads_over_the_gpus():
    from enhancershoebox_b200 import drivers
    assert [drivers.pick_device([r], 8) for r in (1, 2, 8, 9, 100)] == [1, 2, 0, 1, 4]
    assert drivers.pick_device([5], 1) == 0
