import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle.pyoracle import Ref, have_ref, REFERENCE_ROOT
    if not have_ref() and not os.path.exists(REFERENCE_ROOT):
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return Ref()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "pssm.npz"))


@pytest.fixture(scope="session")
def gpu():
    import _pkg
    _pkg.load()
    from mia_b200 import api
    g = api.MiaGpu(0)
    yield g
    g.close()
