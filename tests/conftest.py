import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, '3d-reconstruction-tool_b200')
for p in (ROOT, PKG, os.path.join(PKG, 'src')):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


@pytest.fixture(scope='session')
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from b200dense.engine import DenseEngine
    return DenseEngine('cuda:0')
