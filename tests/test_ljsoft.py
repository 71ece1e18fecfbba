"""a2: the LJsoft table model against the reference generator's own output (golden hashes made by
tests/golden/make_ljsoft_golden.py from /root/reference/ljsoft_potential.py)."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from enhancershoebox_b200 import ljsoft

GOLD = json.load(open(os.path.join(GOLDEN, "ljsoft_sections.json")))


def test_section_list_matches_reference():
    assert [s[0] for s in ljsoft.SECTIONS] == list(GOLD)     # same 55 keywords, same file order
    for key, eps0, s0, a, b, lr in ljsoft.SECTIONS:
        g = GOLD[key]
        assert (float(eps0), float(s0), float(a), float(b), bool(lr)) == (g["eps0"], g["s0"], g["a"], g["b"], g["longrange"]), key


@pytest.mark.parametrize("key", ["CHROMATIN_CHROMATIN", "SER5P_SER5P", "RNP_RNP_ATTRACTION", "ACTIN_SER5P_LONGRANGE",
                                 "PROMOTER_SER5P_STRONG", "RBP_RBP", "CHROMATIN_SER5P"])
def test_section_text_is_byte_identical(key, tmp_path):
    path = tmp_path / "one.table"
    ljsoft.write_table_file(str(path), [key])
    text = path.read_text()
    assert len(text) == GOLD[key]["n_bytes"]
    assert hashlib.sha256(text.encode()).hexdigest() == GOLD[key]["sha256"]
    lines = text.split("\n")
    assert lines[1] == GOLD[key]["header"] == "N\t30000\tRSQ\t0.0001\t3.0"
    assert lines[3] == GOLD[key]["row1"] and lines[3 + 29999] == GOLD[key]["row30000"]


def test_all_sections_row_spot_check():
    # cheap check of every section: rows 1, 4000 and 30000 as the reference prints them
    for key in GOLD:
        rows = list(ljsoft._section_rows(key))
        for idx, name in ((0, "row1"), (3999, "row4000"), (29999, "row30000")):
            r, e, f = rows[idx]
            assert f"{idx + 1}\t{r}\t{e}\t{f}" == GOLD[key][name], (key, name)


def test_table_file_round_trip(tmp_path):
    path = tmp_path / "two.table"
    ljsoft.write_table_file(str(path), ["RBP_RBP", "SER5P_SER5P"])
    secs = ljsoft.read_table_file(str(path))
    assert set(secs) == {"RBP_RBP", "SER5P_SER5P"}
    for key, sec in secs.items():
        r, e, f = ljsoft.section_columns(key)
        assert sec.n == 30000 and sec.rflag == "RSQ" and sec.rlo == 0.0001 and sec.rhi == 3.0
        assert np.array_equal(sec.r, r) and np.array_equal(sec.e, e) and np.array_equal(sec.f, f)
    only = ljsoft.read_table_file(str(path), {"SER5P_SER5P"})
    assert list(only) == ["SER5P_SER5P"]


def test_geometry_of_used_sections():
    # s, rc of the sections the drivers use
    for key, s, rc in (("CHROMATIN_CHROMATIN", 1.0295, 1.0903), ("SER5P_SER5P", 0.3088, 0.8177),
                       ("RNP_RNP_ATTRACTION", 1.0295, 2.7258), ("CHROMATIN_SER5P", 1.3383, 1.4174),
                       ("RBP_RBP", 0.3088, 0.3271), ("PROMOTER_SER5P_STRONG", 0.6692, 1.7718)):
        g = ljsoft.section_geometry(key)
        assert abs(g.s - s) < 5e-5 and abs(g.rc - rc) < 5e-5 and abs(g.c - 0.32) < 1e-15
    # the potential is cut, not force-shifted: F jumps to 0 at rc for the a == b sections
    r, e, f = ljsoft.section_columns("CHROMATIN_CHROMATIN")
    k = np.nonzero(f == 0)[0][0]
    assert abs(r[k - 1] - 1.09025) < 1e-4 and abs(f[k - 1] - 1.6133) < 1e-3 and np.all(f[k:] == 0) and np.all(e[k:] == 0)
