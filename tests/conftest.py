import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = "paper_romualdi_2022_icra_centroidal-mpc-walking_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pkg(sub: str = ""):
    """import the (hyphenated) product package or one of its submodules"""
    return importlib.import_module(PKG + ("." + sub if sub else ""))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as om
    om.build()
    return om.Oracle()


@pytest.fixture(scope="session")
def workloads():
    return pkg("workloads")


def have_ref():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_tmp.so"))
