import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_built():
    import posebyte_b200 as pb
    need = [pb.LIB_PATH, pb.SYNTH_PATH, os.path.join(ROOT, "oracle", "_build", "libposebyte_oracle.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__ as g
        g.build()


@pytest.fixture(scope="session", autouse=True)
def built():
    _ensure_built()


@pytest.fixture(scope="session")
def pb():
    import posebyte_b200
    return posebyte_b200


@pytest.fixture(scope="session")
def orc():
    import oracle_py
    return oracle_py


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible: the product has no CPU path")
    return torch
