import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure; never imported by the product package)."""
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def ops():
    """The product operator layer (sm_100a kernels through the C ABI). Fails loudly if not built."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import lgu_slam_b200
    return lgu_slam_b200.ops
