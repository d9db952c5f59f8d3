import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # keep the in-tree libraries in step with the sources (no-op when they are up to date)
    import __graft_entry__
    __graft_entry__.build()


@pytest.fixture(scope="session")
def orc():
    import oracle_lib
    oracle_lib.build()
    return oracle_lib


@pytest.fixture(scope="session")
def vq():
    import vqengine_b200
    return vqengine_b200


@pytest.fixture(scope="session")
def ctx(vq):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.init()
    c = vq.Context(0)
    yield c
    c.close()
