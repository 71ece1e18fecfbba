"""a1: data-file dialect and the seeded IC generator."""
import os

import numpy as np
import pytest

from enhancershoebox_b200 import datafile as D

REF_IC = "/root/reference/VisitorGeneMaps/input_files/IC_ser5P370_RBP550_0RNP_promoter{p}_bs11.data"


@pytest.mark.parametrize("box,n", [(9, 1050), (10, 1172), (11, 1290), (12, 1400)])
def test_counts(box, n):
    d = D.make_shoebox(box, 3, False, seed=1)
    cfg = D.BOX_TABLE[box]
    assert d.n_atoms == n
    assert d.bonds.shape[0] == cfg["polymer"] - 2 and d.angles.shape[0] == cfg["polymer"] - 4
    np.testing.assert_array_equal(d.box_hi, [box + 2, 9.5, 7.0])
    np.testing.assert_array_equal(d.box_lo, [-box - 2, -9.5, -7.0])
    # anchors: run_single_shoebox.py:651-668 ({1,100,101,300} ... as atom IDs)
    third = cfg["polymer"] // 3
    np.testing.assert_array_equal(d.anchors() + 1, [1, third, third + 1, cfg["polymer"]])
    assert (d.types == D.ENHANCER).sum() == 50 and (d.types == D.INACTIVE_GENE).sum() == 5 and (d.types == D.PROMOTER).sum() == 3
    assert (d.types == D.RBP).sum() == cfg["rbp"] and (d.types == D.SER5P).sum() == cfg["ser5p"]
    g = d.gene_start()
    assert np.all(d.x[g - 3:g + 5, 0] == 0.0)              # gene + promoter beads start at x = 0 (make_input_file.py:230-233)
    assert d.promoter_length() == 3


def test_actin_counts():
    d = D.make_shoebox(11, 3, True, seed=2)
    assert d.n_atoms == 1593                                  # 370 + 3 + 300 + 550 + 370
    assert d.bonds.shape[0] == 368 + 2 and d.angles.shape[0] == 366 + 1
    assert (d.bonds[:, 0] == 2).sum() == 2 and (d.angles[:, 0] == 2).sum() == 1
    assert D.input_file_name(11, 3, True).endswith("IC_ser5P370_actin300_RBP550_0RNP_promoter3_bs11.data")
    assert D.input_file_name(11, 3, False).endswith("IC_ser5P370_RBP550_0RNP_promoter3_bs11.data")


def test_seed_reproducible_and_distinct():
    a, b, c = D.make_shoebox(11, 3, False, seed=5), D.make_shoebox(11, 3, False, seed=5), D.make_shoebox(11, 3, False, seed=6)
    assert np.array_equal(a.x, b.x) and not np.array_equal(a.x, c.x)
    assert np.array_equal(a.types, c.types)
    free = a.x[370:]
    assert np.all(np.abs(free[:, 1]) <= 7.5) and np.all(np.abs(free[:, 2]) <= 5.0)
    assert np.all(a.x[370 + 550:, 0] <= 0.0)                  # Ser5P start in the left half (make_input_file.py:268-272)
    assert np.array_equal(np.round(free, 6), free)


def test_write_read_round_trip(tmp_path):
    d = D.make_shoebox(10, 2, False, seed=3)
    p = tmp_path / "ic.data"
    D.write_data(str(p), d)
    e = D.read_data(str(p))
    for name in ("types", "x", "v", "mol", "bonds", "angles", "box_lo", "box_hi"):
        assert np.array_equal(getattr(d, name), getattr(e, name)), name
    assert e.promoter_length() == 2


def test_read_rejects_inconsistent_counts(tmp_path):
    d = D.make_shoebox(9, 1, False, seed=3)
    p = tmp_path / "ic.data"
    D.write_data(str(p), d)
    txt = p.read_text().replace("1050 atoms", "1051 atoms")
    p.write_text(txt)
    with pytest.raises(ValueError):
        D.read_data(str(p))


@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_bundled_reference_files(p):
    """The production ICs shipped with the reference parse, and the generator reproduces their
    polymer layout (types, bonds, angles, polymer coordinates) exactly."""
    path = REF_IC.format(p=p)
    if not os.path.exists(path):
        pytest.skip("reference checkout not present on this machine")
    ref = D.read_data(path)
    assert (ref.n_atoms, ref.bonds.shape[0], ref.angles.shape[0]) == (1290, 368, 366)
    gen = D.make_shoebox(11, p, False, seed=0)
    assert np.array_equal(ref.types, gen.types) and np.array_equal(ref.bonds, gen.bonds) and np.array_equal(ref.angles, gen.angles)
    assert np.allclose(ref.x[:370], gen.x[:370], atol=0, rtol=0)
    assert ref.gene_start() == gen.gene_start() and ref.promoter_length() == p
    # regulatory region IDs 38-87; promoter3 IDs 244-246, gene 247-251
    assert ref.enhancer_atoms()[0] + 1 == 38 and ref.enhancer_atoms()[-1] + 1 == 87
    if p == 3:
        assert ref.gene_start() + 1 == 247
