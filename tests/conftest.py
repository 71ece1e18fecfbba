import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU and the built libesb.so")


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import esb_oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def pair_tables():
    from enhancershoebox_b200 import protocol as P
    return P.pair_tables("genes_stages")


@pytest.fixture(scope="session")
def sections(pair_tables):
    from enhancershoebox_b200 import ljsoft
    return {k: ljsoft.generated_section(k) for k in sorted({s[0] for s in pair_tables.specs})}


@pytest.fixture(scope="session")
def oracle_lookup(oracle_lib, pair_tables, sections):
    return [oracle_lib.build_lookup_table(sections[k], c) for (k, c) in pair_tables.specs]


@pytest.fixture(scope="session")
def box11():
    from enhancershoebox_b200 import datafile
    return datafile.make_shoebox(11, 3, False, seed=7)


@pytest.fixture(scope="session")
def snapshots():
    return np.load(os.path.join(GOLDEN, "snap_box11.npz"))
