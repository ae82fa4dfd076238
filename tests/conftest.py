import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu on the GPU box')


@pytest.fixture(scope='session')
def golden():
    import numpy as np
    return {name: dict(np.load(os.path.join(GOLDEN, name + '.npz'))) for name in ('sphere_s2', 'spd_s2pp', 'spd_s6pp')}


@pytest.fixture(scope='session')
def mp_truth():
    """60-digit mpmath values of the SPD kernels on the golden inputs (tests/golden/make_mp_truth.py)."""
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, 'spd_truth_mp.npz')))


@pytest.fixture(scope='session')
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda:0')
