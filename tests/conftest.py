import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "tfops_golden.npz"))


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("this test is marked gpu but no CUDA device is visible")
    return torch.device("cuda:0")


@pytest.fixture
def tf32_rounding(cuda):
    """kernel unit tests that pin the tf32 rounding of operand producers: put the library's process-wide switch into
    its default state (an x3 / h3 engine of an earlier test leaves it off)"""
    from monopsr_b200.core.engine import Engine
    Engine.set_library_rounding(cuda, 1)
    return cuda
