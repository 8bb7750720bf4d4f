import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def fx():
    from geos_chem_b200 import grid
    return grid.load_fixture()


@pytest.fixture(scope="session")
def lib():
    """the product library; built in-tree by __graft_entry__.build() (nvcc cross-compiles without a GPU)"""
    from geos_chem_b200 import kpp
    if not os.path.exists(kpp.LIB_PATH):
        from geos_chem_b200 import build
        build.build()
    return kpp.load_library()


@pytest.fixture(scope="session")
def solver(lib):
    from geos_chem_b200 import kpp
    s = kpp.KppSolver("fullchem", device=0, max_cells=1 << 16)
    yield s
    s.close()
