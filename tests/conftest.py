import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'deep-super-resolution_b200')
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden():
    import torch

    def load(name):
        return torch.load(os.path.join(GOLDEN, name))
    return load


def rebuild_z0(fx):
    """Perturbed input of a large fixture (oracle/make_golden_large.py keeps the seed, not the 33.5 MB tensor):
    the CPU draws that follow get_net in the reference flow -- get_noise's uniform (utils/DIP.py:92-96), the
    Downsampler's throw-away Conv2d initialisation (utils/downsampler.py:44) and the closure's noise.normal_()
    (DIP.py:52).  Call right after building the network under torch.manual_seed(fx['seed'])."""
    import torch
    ni = torch.zeros(1, 32, fx['H'], fx['W']).uniform_() * 0.1
    torch.nn.Conv2d(3, 3, kernel_size=4 * fx['factor'], stride=fx['factor'])
    return ni + ni.clone().normal_() * fx['reg_noise_std']
