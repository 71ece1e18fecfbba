"""A chunk run as several scheduling segments (work items) is the same arithmetic as the chunk in one piece."""
import numpy as np
import pytest

from enhancershoebox_b200 import engine as E, protocol as P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_seg", [2, 4, 7])
def test_segmented_chunks_are_bitwise_identical(box11, snapshots, monkeypatch, n_seg):
    x, v, t = snapshots["x500"].astype(np.float64), snapshots["v500"].astype(np.float64), snapshots["t500"].astype(np.int32)
    plans = P.chunk_schedule("genes_stages", n_chunks=503)[500:]
    out = []
    for seg in (1, n_seg):
        monkeypatch.setenv("ESB_SEG", str(seg))
        eng = E.ShoeboxBatch(box11, n_replicas=5, seeds=[3, 4, 5, 6, 7], thresholds=20, activations=100)
        eng.upload_state(x[None], v[None], t[None], replicate=True)
        eng.set_schedule(plans=plans, observe=True)
        eng.run(0, 2)                                   # two chunks in one launch, then one more
        eng.run(2, 1)
        c = eng.counters()
        out.append(eng.state() + (eng.frames(), eng.status(), c["steps"], c["builds"], c["evals"], c["dangerous"], c["listed"], c["pairs"]))
        eng.close()
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)
    assert np.all(out[0][4] == 0) and np.all(out[0][5] == 600)


@pytest.mark.parametrize("cluster", [2, 4, 8])
def test_cluster_of_ctas_per_replica_is_bitwise_identical(box11, snapshots, monkeypatch, cluster):
    """Small batches are stepped by a thread-block cluster per replica (quads split over the CTAs, positions exchanged
    through distributed shared memory): every bead sees the same arithmetic as in the one-CTA kernel."""
    x, v, t = snapshots["x500"].astype(np.float64), snapshots["v500"].astype(np.float64), snapshots["t500"].astype(np.int32)
    plans = P.chunk_schedule("genes_stages", n_chunks=503)[500:]
    out = []
    for size in (1, cluster):
        monkeypatch.setenv("ESB_CLUSTER_SIZE", str(size))
        eng = E.ShoeboxBatch(box11, n_replicas=3, seeds=[3, 4, 5], thresholds=20, activations=100)
        eng.upload_state(x[None], v[None], t[None], replicate=True)
        eng.set_schedule(plans=plans, observe=True)
        eng.run(0, 2)
        eng.run(2, 1)
        c = eng.counters()
        out.append(eng.state() + (eng.frames(), eng.status(), c["steps"], c["builds"], c["evals"], c["dangerous"], c["listed"], c["pairs"]))
        eng.close()
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)
    assert np.all(out[0][4] == 0) and np.all(out[0][5] == 600)


def test_wide_ctas_with_more_replicas_than_sms_are_bitwise_identical(box11, snapshots, monkeypatch):
    """The 32-warp CTA (one per SM) is also the kernel of systems too large for two 16-warp CTAs per SM, whatever
    the batch size: with more replicas than SMs it walks the same ticket ring as the 16-warp kernel."""
    x, v, t = snapshots["x500"].astype(np.float64), snapshots["v500"].astype(np.float64), snapshots["t500"].astype(np.int32)
    plans = P.chunk_schedule("genes_stages", n_chunks=503)[500:]
    R = 170
    out = []
    for wide in (0, 1):
        monkeypatch.setenv("ESB_WIDE", str(wide))
        eng = E.ShoeboxBatch(box11, n_replicas=R, seeds=np.arange(R) + 11, thresholds=20, activations=100)
        eng.upload_state(x[None], v[None], t[None], replicate=True)
        eng.set_schedule(plans=plans, observe=True)
        eng.run(0, 2)
        eng.run(2, 1)
        c = eng.counters()
        out.append(eng.state() + (eng.frames(), eng.status(), c["steps"], c["builds"], c["evals"], c["dangerous"], c["listed"], c["pairs"]))
        eng.close()
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)
    assert np.all(out[0][4] == 0) and np.all(out[0][5] == 600)


def test_16_bit_list_entries_are_bitwise_identical(box11, snapshots, monkeypatch):
    """With at most 32 tables the neighbour lists hold 16-bit entries, two per word, laid out so that every lane sums
    its neighbours in the order of the 32-bit lists: same trajectory, same counters, half the list traffic."""
    x, v, t = snapshots["x500"].astype(np.float64), snapshots["v500"].astype(np.float64), snapshots["t500"].astype(np.int32)
    plans = P.chunk_schedule("genes_stages", n_chunks=503)[500:]
    out = []
    for packed in (0, 1):
        monkeypatch.setenv("ESB_PACK16", str(packed))
        eng = E.ShoeboxBatch(box11, n_replicas=5, seeds=[3, 4, 5, 6, 7], thresholds=20, activations=100)
        eng.upload_state(x[None], v[None], t[None], replicate=True)
        eng.set_schedule(plans=plans, observe=True)
        eng.run(0, 2)
        eng.run(2, 1)
        c = eng.counters()
        out.append(eng.state() + (eng.frames(), eng.status(), c["steps"], c["builds"], c["evals"], c["dangerous"], c["listed"], c["pairs"]))
        eng.close()
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)
    assert np.all(out[0][4] == 0) and np.all(out[0][5] == 600)
