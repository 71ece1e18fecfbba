"""a3-a5: the drivers' command stream replayed as data (SURVEY.md appendix C is the cross-check)."""
import numpy as np
import pytest

from enhancershoebox_b200 import protocol as P


def test_soft_cutoffs_table_c3():
    m = P.soft_cutoffs()
    c_cc = 2 ** (1 / 6)
    exp = {(1, 1): 1.1225, (1, 4): 0.7296, (2, 5): 1.6837, (1, 6): 1.1225, (2, 7): 1.6837, (1, 8): 0.7296, (4, 4): 0.3367,
           (4, 5): 0.7296, (4, 7): 0.7296, (4, 8): 0.3367, (4, 10): 0.7296, (5, 5): 1.1225, (5, 6): 1.6837, (5, 7): 1.1225,
           (5, 8): 0.7296, (5, 11): 1.6837, (6, 7): 1.6837, (6, 8): 0.7296, (6, 9): 1.1225, (7, 7): 1.6837, (7, 8): 0.7296,
           (7, 10): 1.6837, (8, 8): 0.3367, (8, 9): 0.7296, (3, 3): 1.1225, (9, 11): 1.1225}
    for (i, j), v in exp.items():
        assert abs(m[i, j] - v) < 5e-5 and m[i, j] == m[j, i], (i, j)
    assert m[1, 1] == c_cc


def test_table_map_genes_stages_phase_b():
    tm = P.table_map("genes_stages", "Control", "B")
    exp = {(1, 1): ("CHROMATIN_CHROMATIN", 1.2), (1, 4): ("RBP_CHROMATIN", 0.75), (2, 5): ("RNP_CHROMATIN", 1.75),
           (1, 6): ("INDUCED_CHROMATIN", 1.2), (2, 7): ("ACTIVE_CHROMATIN", 1.75), (1, 8): ("CHROMATIN_SER5P", 1.5),
           (2, 8): ("CHROMATIN_SER5P", 1.5), (4, 4): ("RBP_RBP", 0.4), (4, 8): ("RBP_SER5P", 0.4), (4, 11): ("RBP_REGULATORY", 0.75),
           (5, 5): ("RNP_RNP_ATTRACTION", 2.9), (5, 7): ("RNP_ACTIVE", 2.9), (5, 9): ("RNP_REGULATORY", 1.75),
           (5, 10): ("RNP_PROMOTER", 1.75), (5, 11): ("RNP_PROMOTER", 1.75), (6, 7): ("INDUCED_ACTIVE", 1.75),
           (6, 8): ("INDUCED_SER5P", 0.75), (7, 8): ("ACTIVE_SER5P", 1.2), (7, 9): ("ACTIVE_REGULATORY", 1.75),
           (7, 10): ("CHROMATIN_CHROMATIN", 1.2), (8, 8): ("SER5P_SER5P", 0.9), (8, 9): ("SER5P_REGULATORY_WEAK", 1.9),
           (8, 10): ("PROMOTER_SER5P_STRONG", 1.9), (8, 11): ("ACTIVEPROMOTER_SER5P", 1.2), (9, 9): ("CHROMATIN_CHROMATIN", 1.2),
           (6, 6): ("CHROMATIN_CHROMATIN", 1.2), (1, 9): ("CHROMATIN_CHROMATIN", 1.2), (3, 8): ("ACTIN_RBP", 0.4)}
    for k, v in exp.items():
        assert tm[k] == v, k
    assert len(tm) == 66 and all(i <= j for i, j in tm)


def test_table_map_phase_a_ser5p_behaves_like_rbp():
    ta = P.table_map("genes_stages", "Control", "A")
    assert ta[(8, 8)] == ("RBP_SER5P", 0.4) and ta[(1, 8)] == ("RBP_CHROMATIN", 0.75) and ta[(8, 9)] == ("RBP_REGULATORY", 0.75)
    assert ta[(6, 8)] == ("RBP_INDUCED", 0.75) and ta[(7, 8)] == ("RBP_ACTIVE", 0.75) and ta[(8, 11)] == ("RBP_REGULATORY", 0.75)


def test_table_map_shoebox_actin_rows():
    tb = P.table_map("shoebox", "Control", "B")
    exp = {(1, 3): ("ACTIN_CHROMATIN", 1.2), (3, 3): ("ACTIN_ACTIN", 0.4), (3, 4): ("ACTIN_RBP", 0.4), (3, 5): ("RNP_ACTIN", 1.9),
           (3, 6): ("ACTIN_INDUCED", 1.2), (3, 7): ("ACTIN_ACTIVE", 1.2), (3, 8): ("ACTIN_SER5P", 0.9), (3, 9): ("ACTIN_REGULATORY", 0.75),
           (3, 10): ("ACTIN_PROMOTER", 0.75), (3, 11): ("ACTIN_PROMOTER", 0.75)}
    for k, v in exp.items():
        assert tb[k] == v, k
    assert P.table_map("shoebox", "SwinA_nopol", "B")[(3, 3)] == ("ACTIN_ACTIN_STRONGATTRACTION", 0.9)
    assert P.table_map("shoebox", "noActinSer5P", "B")[(3, 8)] == ("ACTIN_SER5P_NOATTRACTION", 0.4)
    assert P.table_map("shoebox", "Control", "A")[(3, 8)] == ("ACTIN_RBP", 0.4)
    assert P.table_map("shoebox", "Hexanediol", "T")[(8, 8)] == ("RBP_RBP", 0.4)
    assert P.table_map("shoebox", "JQ-1", "T")[(8, 9)] == ("RBP_REGULATORY", 0.75)


def test_distinct_tables():
    pt = P.pair_tables("genes_stages")
    assert len(pt.specs) == 29                                  # 29 distinct (KEY, cut)
    for ph in ("A", "B"):
        m = pt.maps[ph]
        assert np.array_equal(m, m.T) and np.all(m[1:, 1:] >= 0)


def test_chunk_schedule():
    pl = P.chunk_schedule("genes_stages", n_chunks=20)
    assert [p.pair for p in pl] == ["soft"] * 9 + ["A"] * 5 + ["B"] * 6
    assert all(p.ramp == (0, 2000) and p.adapt_soft and p.adapt_bond1 for p in pl[:9])
    assert all(p.ramp is None and not p.adapt_soft and not p.adapt_bond1 for p in pl[9:])
    assert not any(p.adapt_bond2 for p in pl)
    sb = P.chunk_schedule("shoebox", n_chunks=12)
    assert all(p.adapt_bond2 for p in sb)                        # fix s3 is never unfixed (run_single_shoebox.py:638-639, 794-795)
    assert len(P.chunk_schedule("shoebox", "Hexanediol")) == 2000 + 300
    assert P.chunk_schedule("shoebox", "Hexanediol")[2000].pair == "T" and P.chunk_schedule("shoebox", "Flavopiridol")[2500].pair == "B"
    assert P.DELT == 6.0 and P.T_RUN == 200 and P.N_RUNS == 2000
