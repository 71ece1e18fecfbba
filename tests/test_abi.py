"""(b) the C-ABI boundary on a machine without a GPU: the library loads, exports every symbol
include/esb.h declares, and refuses to compute without a device (no silent CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, has_gpu
from enhancershoebox_b200 import build as B
from enhancershoebox_b200 import engine as E


@pytest.fixture(scope="module")
def lib():
    B.build()
    return E.load_library()


def _declared():
    text = open(os.path.join(ROOT, "include", "esb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(esb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/esb.h but not exported by libesb.so"
    assert sorted(E.ABI_SYMBOLS) == names            # the Python binding covers the whole header, no more


def test_version_and_error_string(lib):
    assert lib.esb_version() >= 100
    assert isinstance(lib.esb_last_error(), bytes)


def test_table_builder_is_host_only_and_matches_oracle(lib, oracle_lib, sections):
    """esb_table_from_section (product) and the oracle's builder are independent transcriptions of
    the same published algorithm and must agree to the last bit."""
    for key, cut in (("CHROMATIN_CHROMATIN", 1.2), ("SER5P_SER5P", 0.9), ("RNP_RNP_ATTRACTION", 2.9), ("RBP_RBP", 0.4)):
        e1, f1, i1, d1 = E.build_lookup(sections[key], cut)
        e2, f2, i2, d2 = oracle_lib.build_lookup_table(sections[key], cut)
        assert np.array_equal(e1, e2) and np.array_equal(f1, f2) and (i1, d1) == (i2, d2)
    with pytest.raises(E.EsbError, match="cutoff outside"):
        E.build_lookup(sections["RBP_RBP"], 3.5)


@pytest.mark.skipif(has_gpu(), reason="this check is for GPU-less machines")
def test_no_cpu_fallback(lib, box11):
    h = C.c_void_p()
    rc = lib.esb_create(0, 1, 100, 11, C.byref(h))
    assert rc == -3 and b"no CPU fallback" in lib.esb_last_error()        # ESB_ENODEV
    with pytest.raises(E.EsbError):
        E.ShoeboxBatch(box11, n_replicas=1)
    out = np.zeros((1, 4))
    fr = np.zeros((5, 200, 9))
    rc = lib.esb_aggregate_grouped(0, fr.ctypes.data_as(C.c_void_p), 0, 5, 200, 5, 250.0, 150,
                                   out.ctypes.data_as(C.POINTER(C.c_double)), None)
    assert rc == -3


def test_product_never_imports_the_oracle():
    """No import, include, dlopen or symbol of oracle/ anywhere in the product package."""
    pkg = os.path.join(ROOT, "enhancershoebox_b200")
    bad = re.compile(r"(^\s*(from|import)\s+oracle\b|from\s+\.+\s*oracle|oracle/|libesb_oracle|esbo_|esb_oracle)", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not bad.search(src), os.path.join(dirpath, f)
