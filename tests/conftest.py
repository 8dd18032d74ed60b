import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def cuda_ops():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from refvsr_b200.lib import CudaOps
    return CudaOps()          # raises if librefvsr_b200.so is missing: no silent fallback


@pytest.fixture(scope='session')
def oracle_ops():
    from oracle.oracle_ops import OracleOps
    return OracleOps()
